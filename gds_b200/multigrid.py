"""Multigrid preconditioner for the node-domain solves of the path: the pressure problem of `set_evolution(lhs=...)`
(gds/fds.py:97-113,317-328, used by examples/fluid.py:24-38) and the Leray projection (gds/gds.py:249,414-419).

The reference hands these to a general convex solver / a dense pseudo-inverse.  On the device they are conjugate-gradient
solves whose operator is a (Dirichlet-masked, possibly scaled) graph Laplacian; unpreconditioned, a lattice of V nodes needs
O(sqrt(V)) x 10 iterations (1.5e4 at V = 8e6).  This module builds, ONCE per (graph, Dirichlet set), a smoothed-aggregation
hierarchy of the masked graph Laplacian M0 on the host (scipy sparse products; the aggregation sweep is native code,
`gds_mg_aggregate`) and uploads it through the C-ABI (`gds_mg_*`); `gds_pcg_solve` then applies one symmetric V-cycle per
iteration.  Conjugate gradients are invariant under a positive scaling of the preconditioner, so the hierarchy of M0 serves
every operator c * Z_D(L0 .) whatever c; the operator of the solve itself stays the caller's pass, so the answer satisfies
the caller's system to the requested tolerance whatever the quality of the hierarchy (only the iteration count depends on it).

Everything here is in DEVICE row order (graph.LoweredGraph.p2i) because that is the order of the vectors it is applied to.
"""
import ctypes as C
import weakref

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from ._lib import check, lib, ptr

THETA = 0.5           # strength of connection: |a_ij| >= THETA * (largest off-diagonal magnitude of row i)
COARSE_SIZE = 400     # levels at most this large are solved with a dense pseudo-inverse
MAX_LEVELS = 12


def masked_laplacian(low, dirichlet_rows=()):
    """M0 = B1 diag(w) B1^T (= -L0, gds.py:105-109,136-137) in device row order, Dirichlet rows and columns replaced by
    identity rows (they carry no unknown: the solves keep those entries at zero)."""
    V = low.V
    u, v = np.asarray(low.edge_u, dtype=np.int64), np.asarray(low.edge_v, dtype=np.int64)
    if low.p2i is not None:
        u, v = low.p2i[u], low.p2i[v]
    w = np.asarray(low.weights, dtype=np.float64)
    keep = u != v                                  # a self-loop contributes nothing to B1 diag(w) B1^T's off-diagonal
    free = np.ones(V, dtype=bool)
    D = np.asarray(low.rows_to_dev(L.NODES, np.asarray(dirichlet_rows, dtype=np.int64)), dtype=np.int64)
    free[D] = False
    deg = np.bincount(u[keep], weights=w[keep], minlength=V) + np.bincount(v[keep], weights=w[keep], minlength=V)
    inner = keep & free[u] & free[v]
    ui, vi, wi = u[inner], v[inner], w[inner]
    diag = np.where(free, deg, 1.0)
    rows = np.concatenate([ui, vi, np.arange(V, dtype=np.int64)])
    cols = np.concatenate([vi, ui, np.arange(V, dtype=np.int64)])
    vals = np.concatenate([-wi, -wi, diag])
    A = sp.csr_matrix((vals, (rows, cols)), shape=(V, V))
    A.sum_duplicates()
    A.sort_indices()
    return A


def aggregate(A, theta=None):
    """aggregate id of every row (-1: no strong neighbour, no coarse representation) and the number of aggregates"""
    theta = THETA if theta is None else theta
    A = sp.csr_matrix(A)
    n = A.shape[0]
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    agg = np.empty(n, dtype=np.int32)
    na = C.c_int64()
    check(lib.gds_mg_aggregate(n, ptr(rp), ptr(col), ptr(val), float(theta), ptr(agg), C.byref(na)))
    return agg, int(na.value)


def jacobi_weights(A, power_iterations=30):
    """omega / a_ii with omega = 4 / (3 rho), rho an estimate of the spectral radius of D^-1 A: a few power iterations
    (x 1.1), never above the Gershgorin bound.  The damped sweep I - omega D^-1 A contracts as long as the true radius is
    below 1.5 rho, so the estimate has a wide margin; the bound alone (2 for a Laplacian) over-damps the coarse levels,
    whose radius is about 1.5, and costs a third more iterations."""
    d = A.diagonal()
    dinv = np.zeros_like(d)
    nz = d != 0
    dinv[nz] = 1.0 / d[nz]
    rowsum = np.asarray(abs(A).sum(axis=1)).ravel()
    rho = float(np.max(np.abs(dinv) * rowsum)) if A.shape[0] else 1.0
    if A.shape[0] > 1 and power_iterations > 0:
        x = np.random.default_rng(0).standard_normal(A.shape[0])
        lam = 0.0
        for _ in range(power_iterations):
            y = dinv * (A @ x)
            ny = float(np.linalg.norm(y))
            if ny == 0.0:
                break
            lam = ny / float(np.linalg.norm(x))
            x = y / ny
        if lam > 0.0:
            rho = min(rho, 1.1 * lam)
    omega = 4.0 / (3.0 * max(rho, 1e-300))
    return omega * dinv, omega, rho


def build_levels(A, coarse_size=None, max_levels=MAX_LEVELS, theta=None):
    """smoothed aggregation: tentative piecewise-constant prolongation T over the aggregates, P = (I - omega D^-1 A) T,
    Galerkin coarse matrix P^T A P; the last level carries the dense pseudo-inverse of its matrix (the Laplacian of a graph
    without Dirichlet rows is singular: constants per component)"""
    coarse_size = COARSE_SIZE if coarse_size is None else coarse_size
    theta = THETA if theta is None else theta
    A = sp.csr_matrix(A, dtype=np.float64)
    levels = []
    while True:
        n = A.shape[0]
        w, omega, rho = jacobi_weights(A)
        lev = {'A': A, 'w': w, 'omega': omega, 'rho': rho}
        levels.append(lev)
        last = n <= coarse_size or len(levels) >= max_levels
        if not last:
            agg, na = aggregate(A, theta)
            last = na == 0 or na >= n          # nothing to coarsen (no strong connections at all)
        if last:
            if n > 4096:
                raise ValueError(f'multigrid: coarsening stalled at {n} rows (too large for a dense coarsest solve)')
            lev['pinv'] = np.linalg.pinv(A.toarray(), rcond=1e-10, hermitian=True)
            break
        rows = np.nonzero(agg >= 0)[0]
        T = sp.csr_matrix((np.ones(rows.size), (rows, agg[rows].astype(np.int64))), shape=(n, na))
        P = (T - sp.diags(w) @ (A @ T)).tocsr()
        P.sort_indices()
        R = P.T.tocsr()
        R.sort_indices()
        lev['P'], lev['R'] = P, R
        A = (R @ (A @ P)).tocsr()
        A.sum_duplicates()
        A.eliminate_zeros()
        A.sort_indices()
    return levels


def vcycle_host(levels, b, li=0, pre=1, post=1):
    """the V-cycle the device runs, in numpy (tests compare gds_mg_apply with it)"""
    lev = levels[li]
    if 'pinv' in lev:
        return lev['pinv'] @ b
    A, w = lev['A'], lev['w']
    x = w * b
    for _ in range(pre - 1):
        x = x + w * (b - A @ x)
    x = x + lev['P'] @ vcycle_host(levels, lev['R'] @ (b - A @ x), li + 1, pre, post)
    for _ in range(post):
        x = x + w * (b - A @ x)
    return x


def _csr_args(M):
    M = sp.csr_matrix(M)
    rp = np.ascontiguousarray(M.indptr, dtype=np.int64)
    col = np.ascontiguousarray(M.indices, dtype=np.int32)
    val = np.ascontiguousarray(M.data, dtype=np.float64)
    return rp, col, val


class Multigrid:
    """device copy of a hierarchy (handle of gds_mg_*)"""

    def __init__(self, ctx, levels, sweeps=(1, 1)):
        self.ctx = ctx
        self.h = L.c_vp()
        self.n = levels[0]['A'].shape[0]
        self.sizes = [lev['A'].shape[0] for lev in levels]
        self.nnz = [int(lev['A'].nnz) for lev in levels]
        check(lib.gds_mg_create(ctx.h, C.byref(self.h)))
        self._fin = weakref.finalize(self, lib.gds_mg_destroy, self.h)
        for lev in levels:
            if 'pinv' in lev:
                a = _csr_args(sp.csr_matrix(np.ascontiguousarray(lev['pinv'])))
                check(lib.gds_mg_add_level(self.h, lev['A'].shape[0], ptr(a[0]), ptr(a[1]), ptr(a[2]), None, 0,
                                           None, None, None, None, None, None))
            else:
                a, p, r = _csr_args(lev['A']), _csr_args(lev['P']), _csr_args(lev['R'])
                w = np.ascontiguousarray(lev['w'], dtype=np.float64)
                check(lib.gds_mg_add_level(self.h, lev['A'].shape[0], ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(w),
                                           lev['P'].shape[1], ptr(p[0]), ptr(p[1]), ptr(p[2]), ptr(r[0]), ptr(r[1]), ptr(r[2])))
        check(lib.gds_mg_set_sweeps(self.h, int(sweeps[0]), int(sweeps[1])))

    @classmethod
    def for_graph(cls, low, dirichlet_rows=(), sweeps=(1, 1)):
        """hierarchy of the masked Laplacian of a lowered graph, cached on the graph per Dirichlet set"""
        cache = low.__dict__.setdefault('_mg_cache', {})
        key = (np.asarray(dirichlet_rows, dtype=np.int64).tobytes(), tuple(sweeps))
        if key not in cache:
            if len(cache) >= 4:
                cache.clear()
            cache[key] = cls(low.ctx, build_levels(masked_laplacian(low, dirichlet_rows)), sweeps)
        return cache[key]

    def info(self):
        nl, nb, k = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib.gds_mg_info(self.h, C.byref(nl), C.byref(nb), C.byref(k)))
        return {'levels': nl.value, 'device_bytes': nb.value, 'launches_per_cycle': k.value, 'sizes': self.sizes,
                'operator_complexity': sum(self.nnz) / max(1, self.nnz[0])}

    def apply(self, r, z):
        """z = M^-1 r on two device buffers (one V-cycle)"""
        check(lib.gds_mg_apply(self.h, r.h, z.h))
