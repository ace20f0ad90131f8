"""Host side of the multigrid preconditioner (gds_b200/multigrid.py): the native aggregation sweep, the smoothed-aggregation
hierarchy of the masked graph Laplacian and the V-cycle as a conjugate-gradient preconditioner -- all on the CPU (the numpy
V-cycle here is the model the device V-cycle is compared with in tests/test_multigrid_gpu.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
import gds_b200
from gds_b200 import multigrid as mg
from gds_b200.graph import LoweredGraph


def host_ctx():
    return gds_b200.Context(-1)


def lowered(G):
    return LoweredGraph(G, host_ctx())


def pcg(A, b, M, rtol=1e-10, maxit=500):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, maxit + 1):
        ap = A @ p
        alpha = rz / (p @ ap)
        x += alpha * p
        r -= alpha * ap
        if r @ r <= rtol * rtol * bb:
            return x, it
        z = M(r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
    return x, maxit


GRAPHS = {
    'hex': lambda: cases.g_hex(40, 40),
    'tri_weighted': lambda: cases.add_weights(cases.g_tri(60, 60), 3),
    'grid': lambda: cases.g_grid(70, 70),
    'rgg': lambda: cases.g_rgg(3000, 0.04, 5),
}


def test_masked_laplacian_is_minus_L0_with_identity_dirichlet_rows():
    G = cases.add_weights(cases.g_hex(5, 6), 1)
    low = lowered(G)
    A = mg.masked_laplacian(low).toarray()
    nodes = {v: i for i, v in enumerate(G.nodes())}
    want = np.zeros_like(A)
    for u, v, d in G.edges(data=True):
        i, j, w = nodes[u], nodes[v], d['weight']
        want[i, i] += w; want[j, j] += w; want[i, j] -= w; want[j, i] -= w
    perm = low.p2i if low.p2i is not None else np.arange(len(nodes))
    P = np.zeros_like(A); P[perm, np.arange(len(nodes))] = 1.0          # device row of public node q is p2i[q]
    assert np.allclose(A, P @ want @ P.T, rtol=0, atol=1e-14)
    D = [0, 7, 11]
    Am = mg.masked_laplacian(low, D).toarray()
    Dd = low.rows_to_dev(gds_b200._lib.NODES, np.array(D))
    free = np.setdiff1d(np.arange(len(nodes)), Dd)
    assert np.array_equal(Am[np.ix_(free, free)], A[np.ix_(free, free)])
    for d in Dd:
        row = np.zeros(len(nodes)); row[d] = 1.0
        assert np.array_equal(Am[d], row) and np.array_equal(Am[:, d], row)


@pytest.mark.parametrize('gname', list(GRAPHS))
def test_aggregates_cover_connected_rows_and_are_connected(gname):
    low = lowered(GRAPHS[gname]())
    A = mg.masked_laplacian(low, dirichlet_rows=range(0, low.V, 97))
    agg, na = mg.aggregate(A)
    assert agg.shape == (low.V,) and agg.max() == na - 1 and na < low.V / 2
    offdiag = A - sp.diags(A.diagonal())
    has_nbr = np.asarray(abs(offdiag).sum(axis=1)).ravel() > 0
    assert np.array_equal(agg >= 0, has_nbr)            # identity (Dirichlet) rows and isolated nodes stay outside
    assert np.array_equal(np.unique(agg[agg >= 0]), np.arange(na))
    # every aggregate is a connected piece of the graph
    C = offdiag.tocoo()
    same = (agg[C.row] == agg[C.col]) & (agg[C.row] >= 0)
    H = sp.coo_matrix((np.ones(same.sum()), (C.row[same], C.col[same])), shape=A.shape)
    ncomp, lab = sp.csgraph.connected_components(H, directed=False)
    inside = agg >= 0
    assert len(set(zip(agg[inside], lab[inside]))) == na
    # a pure function of the matrix
    agg2, na2 = mg.aggregate(A)
    assert na2 == na and np.array_equal(agg, agg2)


def test_aggregate_rejects_bad_input():
    A = sp.csr_matrix(np.array([[2., -1.], [-1., 2.]]))
    rp, col, val = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.copy()
    agg, na = np.empty(2, np.int32), gds_b200._lib.c_i64()
    L = gds_b200._lib
    import ctypes as C
    bad = col.copy(); bad[0] = 5
    assert L.lib.gds_mg_aggregate(2, L.ptr(rp), L.ptr(bad), L.ptr(val), 0.5, L.ptr(agg), C.byref(na)) != 0
    assert b'out of range' in L.lib.gds_last_error()
    assert L.lib.gds_mg_aggregate(2, L.ptr(rp), L.ptr(col), L.ptr(val), -1.0, L.ptr(agg), C.byref(na)) != 0
    assert L.lib.gds_mg_aggregate(2, L.ptr(rp), L.ptr(col), L.ptr(val), 1.5, L.ptr(agg), C.byref(na)) != 0
    # an empty matrix and a matrix without off-diagonal entries have no aggregates
    e = sp.identity(3, format='csr')
    agg3, na3 = mg.aggregate(e)
    assert na3 == 0 and np.all(agg3 == -1)


@pytest.mark.parametrize('gname', list(GRAPHS))
def test_hierarchy_is_galerkin_and_the_vcycle_is_a_spd_preconditioner(gname):
    low = lowered(GRAPHS[gname]())
    dirichlet = () if gname in ('hex', 'rgg') else np.arange(0, low.V, 53)        # singular and definite problems
    A = mg.masked_laplacian(low, dirichlet)
    levels = mg.build_levels(A, coarse_size=200)
    assert len(levels) >= 2 and 'pinv' in levels[-1] and levels[-1]['A'].shape[0] <= 200
    for a, b in zip(levels[:-1], levels[1:]):
        assert abs(a['R'] - a['P'].T).max() == 0
        G = (a['P'].T @ a['A'] @ a['P'])
        assert abs(G - b['A']).max() <= 1e-12 * abs(G).max()
        assert abs(b['A'] - b['A'].T).max() <= 1e-12 * abs(G).max()
        assert a['rho'] <= 2.0 + 1e-9 or a is not levels[0]          # never above a Laplacian's Gershgorin bound of 2
    assert sum(l['A'].nnz for l in levels) / levels[0]['A'].nnz < 2.2        # operator complexity
    rng = np.random.default_rng(0)
    # symmetric: <M^-1 a, b> = <a, M^-1 b>
    a, b = rng.standard_normal(low.V), rng.standard_normal(low.V)
    Ma, Mb = mg.vcycle_host(levels, a), mg.vcycle_host(levels, b)
    assert abs(Ma @ b - a @ Mb) <= 1e-10 * (np.linalg.norm(Ma) * np.linalg.norm(b))
    assert a @ Ma > 0 and b @ Mb > 0
    # right-hand side in the range: zero on Dirichlet rows, zero sum per component when singular
    rhs = rng.standard_normal(low.V)
    rhs[low.rows_to_dev(gds_b200._lib.NODES, np.asarray(dirichlet, dtype=np.int64))] = 0.0
    if len(dirichlet) == 0:
        ncomp, lab = low.component_labels()
        rhs -= (np.bincount(lab, weights=rhs, minlength=ncomp) / np.bincount(lab, minlength=ncomp))[lab]
    x, it = pcg(A, rhs, lambda r: mg.vcycle_host(levels, r), rtol=1e-10)
    x0, it0 = pcg(A, rhs, lambda r: r, rtol=1e-10, maxit=5000)
    assert np.linalg.norm(A @ x - rhs) <= 2e-10 * np.linalg.norm(rhs)
    assert it <= 30 and it0 >= 4 * it, (it, it0)
    # two sweeps each side: fewer iterations, still symmetric
    x2, it2 = pcg(A, rhs, lambda r: mg.vcycle_host(levels, r, pre=2, post=2), rtol=1e-10)
    assert it2 <= it


def test_small_graphs_are_a_single_dense_level():
    low = lowered(cases.g_grid(6, 5))
    levels = mg.build_levels(mg.masked_laplacian(low))
    assert len(levels) == 1 and 'pinv' in levels[0]
    b = np.random.default_rng(1).standard_normal(low.V); b -= b.mean()
    x = mg.vcycle_host(levels, b)
    assert np.linalg.norm(levels[0]['A'] @ x - b) <= 1e-10 * np.linalg.norm(b)
