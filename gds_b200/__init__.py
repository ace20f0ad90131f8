"""gds_b200 -- B200-native (sm_100a) stepping core behind the asrvsn/gds field API.

    import gds_b200 as gds
    T = gds.node_gds(G)
    T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2, integrator=gds.Integrators.rk4)
    T.set_initial(y0=...); T.set_constraints(dirichlet=...)
    T.step(0.1)

All arithmetic runs in gds_b200/lib/libgds_b200.so (built by `make`); importing without it raises -- there is no
CPU fallback.
"""
from ._lib import GdsError, LIB_PATH  # noqa: F401  (raises ImportError if the CUDA library is not built)
from .types import (GraphDomain, IterationMode, Integrators, Steppable, Observable, PointObservable,  # noqa: F401
                    VectorObservable)
from .utils import bidict, fun_ary, dict_fun, combine_bcs, const_edge_bc, zero_edge_bc  # noqa: F401
from .device import Context, Buffer, Mask, DeviceGraph  # noqa: F401
from .graph import LoweredGraph, NativeGraph, distribute  # noqa: F401
from .partition import Partition, HaloExchange, split_bounds  # noqa: F401
from .engine import StepEngine, step_table  # noqa: F401
from .lattices import (set_grid_graph_layout, square_lattice, triangular_lattice, hexagonal_lattice,  # noqa: F401
                       get_planar_boundary, embedded_faces, get_edge_weights, set_edge_weights, get_edge_lengths,
                       edge_domain, remove_edges, remove_pos, degree_matrix)
from .fields import (fds, gds, GraphObservable, node_gds, edge_gds, simplex_gds, face_gds, coupled_fds, couple,  # noqa: F401
                     System, BoundarySpec)

__version__ = '0.1.0'
