"""Loader for the UNMODIFIED reference (asrvsn/gds at /root/reference) -- TEST INFRASTRUCTURE ONLY.

Used only by tests/golden/make_golden.py (run in the build container, where /root/reference
exists) to generate golden vectors.  Nothing in the product path, the `-m gpu` tests, smoke()
or bench.py imports this module: /root/reference does not exist on the GPU box.

`import gds` fails here because non-numerical dependencies (cvxpy, shortuuid, hickle, shapely,
trimesh, matplotlib, bokeh, colorcet, selenium, ujson, tornado ...) are absent and networkx 3
removed `to_scipy_sparse_matrix` (used at gds/utils/graph/common.py:6).  We install import-time
stand-ins for those modules; the numerical path (gds/gds.py, gds/fds.py) then runs unmodified on
numpy/scipy/networkx.  `shapely` gets a real point-in-closed-polygon stand-in because
embedded_faces() (gds/utils/graph/generators.py:498-506) uses Polygon.intersects(Point) to pick
the outer face; a truthy mock would silently delete face 0.
"""
import importlib.abc
import importlib.machinery
import sys
import types
import uuid
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"

_ABSENT = ("cvxpy", "shortuuid", "hickle", "trimesh", "matplotlib", "bokeh", "colorcet",
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
    import os
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present (only available in the build container)")
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
    import shortuuid  # the mock
    shortuuid.uuid = lambda: uuid.uuid4().hex
    sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gds as ref_gds
    _loaded = ref_gds
    return ref_gds
