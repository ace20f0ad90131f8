"""Loader for the UNMODIFIED reference (asrvsn/gds) -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python.  It is imported from /root/reference where that tree is mounted (the build
container), else from oracle/_ref/gds_reference.zip -- its `gds` package, unmodified, packed by oracle/build_ref.py at
build() time and imported through zipimport (git-ignored, so it stays out of history, but not gpurun-ignored, so it travels to the GPU box
like a built .so).  Used by tests/golden/make_golden.py (golden vectors), tests/test_golden_fresh_cpu.py and
the CPU legs of bench.py (`--impl reference`, `cpu_baseline`: the reference itself timed on the host cores).
Nothing in the product path (gds_b200/) imports this module.

`import gds` fails here because non-numerical dependencies (shortuuid, hickle, shapely,
trimesh, matplotlib, bokeh, colorcet, selenium, ujson, tornado ...) are absent and networkx 3
removed `to_scipy_sparse_matrix` (used at gds/utils/graph/common.py:6).  We install import-time
stand-ins for those modules; the numerical path (gds/gds.py, gds/fds.py) then runs unmodified on
numpy/scipy/networkx.  `shapely` gets a real point-in-closed-polygon stand-in because
embedded_faces() (gds/utils/graph/generators.py:498-506) uses Polygon.intersects(Point) to pick
the outer face; a truthy mock would silently delete face 0.  `cvxpy` (absent as well) gets oracle/cvx_shim.py: the
affine-expression subset the reference's `lhs=` path touches, with the quadratic programme solved exactly by dense least
squares -- so that path of the reference runs too (see that module's header for what this does and does not pin).
"""
import importlib.abc
import importlib.machinery
import sys
import types
import uuid
from unittest.mock import MagicMock

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_ARCHIVE = os.path.join(_HERE, "_ref", "gds_reference.zip")


def reference_root():
    """sys.path entry holding the reference's `gds` package -- the mounted tree, else the archive -- or None"""
    if os.path.isfile("/root/reference/gds/gds.py"):
        return "/root/reference"
    if os.path.isfile(_ARCHIVE):
        return _ARCHIVE
    return None


REFERENCE_ROOT = reference_root()

_ABSENT = ("shortuuid", "hickle", "trimesh", "matplotlib", "bokeh", "colorcet",
           "selenium", "ujson", "tornado", "seaborn", "statsmodels", "pyunicorn", "zmq")


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _MockFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        top = fullname.split(".")[0]
        if top in _ABSENT:
            return importlib.machinery.ModuleSpec(fullname, _MockLoader(), is_package=True)
        return None


def _make_shapely():
    """Closed-polygon containment stand-in for shapely.geometry.{Point,Polygon}."""
    shapely = types.ModuleType("shapely")
    geometry = types.ModuleType("shapely.geometry")

    class Point:
        def __init__(self, x, y):
            self.x, self.y = float(x), float(y)

    class Polygon:
        def __init__(self, pts):
            self.pts = [(float(p[0]), float(p[1])) for p in pts]

        def intersects(self, p):
            # boundary counts as intersecting (shapely semantics for closed sets)
            x, y = p.x, p.y
            n = len(self.pts)
            inside = False
            for i in range(n):
                x1, y1 = self.pts[i]
                x2, y2 = self.pts[(i + 1) % n]
                # on-segment test
                cross = (x2 - x1) * (y - y1) - (y2 - y1) * (x - x1)
                if abs(cross) <= 1e-12 * max(1.0, abs(x2 - x1) + abs(y2 - y1)):
                    if min(x1, x2) - 1e-12 <= x <= max(x1, x2) + 1e-12 and \
                       min(y1, y2) - 1e-12 <= y <= max(y1, y2) + 1e-12:
                        return True
                if (y1 > y) != (y2 > y):
                    xin = x1 + (y - y1) * (x2 - x1) / (y2 - y1)
                    if xin > x:
                        inside = not inside
            return inside

    geometry.Point = Point
    geometry.Polygon = Polygon
    shapely.geometry = geometry
    return shapely, geometry


_loaded = None


def load_reference():
    """Return the reference's `gds` package, imported unmodified from /root/reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present: neither /root/reference nor oracle/_ref (python oracle/build_ref.py)")
    import networkx as nx
    if not hasattr(nx, "to_scipy_sparse_matrix"):
        nx.to_scipy_sparse_matrix = nx.to_scipy_sparse_array
    import numpy as np
    if not hasattr(np, "core"):
        import numpy._core as _c
        np.core = _c
    sys.meta_path.insert(0, _MockFinder())
    shapely, geometry = _make_shapely()
    sys.modules["shapely"] = shapely
    sys.modules["shapely.geometry"] = geometry
    try:            # CVXPY itself where it exists; else the least-squares stand-in (oracle/cvx_shim.py) for the lhs= path
        import cvxpy  # noqa: F401
    except ImportError:
        from . import cvx_shim
        sys.modules["cvxpy"] = cvx_shim
    import shortuuid  # the mock
    shortuuid.uuid = lambda: uuid.uuid4().hex
    sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gds as ref_gds
    _loaded = ref_gds
    return ref_gds
