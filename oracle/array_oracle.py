"""Array-form CPU oracle for graphs too large for per-point Python index maps -- TEST INFRASTRUCTURE ONLY.

oracle/gds_oracle.py restates the reference object by object (nx.Graph, dict index maps, per-point callables) and is
pinned against golden vectors of the unmodified reference; like the reference it cannot hold 10^7..10^8 rows (a Python
dict entry per row, a Python call per boundary point).  This module restates the SAME formulas over plain arrays -- edge
list in G.edges() order, weights, faces as a CSR list of node indices, boundary rows as index arrays -- so that the CUDA
path can be checked against a CPU result at the sizes bench.py runs (BASELINE configs at full size).  Every function
cites the reference lines it follows; tests/test_array_oracle_cpu.py pins it against gds_oracle.py
(hence, transitively, against the reference's golden vectors) on small graphs, operator by operator and step by step.

Only tests/ imports this module; the product path never does.
"""
import numpy as np
import scipy.sparse as sp

from .gds_oracle import RK4, Euler  # the fixed-step solver objects (protocol of gds/ode/midpoint.py:10-32)


def incidence(V, eu, ev, w=None):
    """B1 . diag(sqrt w), CSR V x E: B[u,e] = -sqrt(w_e), B[v,e] = +sqrt(w_e) for (u, v) = edges_i[e]   (gds/gds.py:109)"""
    E = len(eu)
    omega = np.ones(E) if w is None else np.sqrt(np.asarray(w, dtype=float))
    ar = np.arange(E)
    return sp.csr_matrix((np.concatenate([-omega, omega]), (np.concatenate([eu, ev]), np.concatenate([ar, ar]))),
                         shape=(V, E))


def curl_matrix(V, eu, ev, w, face_ptr, face_nodes):
    """B2 (F x E), gds/gds.py:227-239: for a face (n_0 .. n_{k-1}) the edge between consecutive nodes (n_{i-1}, n_i),
    (n_{k-1}, n_0) gets +sqrt(w_e) when its stored orientation follows the traversal and -sqrt(w_e) otherwise.  (The
    reference's test is "both endpoints on the face", which would also count a chord; faces of the planar lattices have
    none -- SURVEY.md App. A.2.)"""
    E, F = len(eu), len(face_ptr) - 1
    if F == 0:
        return sp.csr_matrix((1, E))                                    # gds.py:239
    omega = np.ones(E) if w is None else np.sqrt(np.asarray(w, dtype=float))
    eu, ev = np.asarray(eu, dtype=np.int64), np.asarray(ev, dtype=np.int64)
    key = np.minimum(eu, ev) * V + np.maximum(eu, ev)
    order = np.argsort(key, kind='stable')
    skey = key[order]
    fp, fn = np.asarray(face_ptr, dtype=np.int64), np.asarray(face_nodes, dtype=np.int64)
    sizes = fp[1:] - fp[:-1]
    face_of = np.repeat(np.arange(F), sizes)
    pos = np.arange(fn.size) - np.repeat(fp[:-1], sizes)
    prev = np.where(pos == 0, np.repeat(fp[1:] - 1, sizes), np.arange(fn.size) - 1)   # (n_{i-1}, n_i), wrapping to n_{k-1}
    a, b = fn[prev], fn[np.arange(fn.size)]
    k = np.minimum(a, b) * V + np.maximum(a, b)
    at = np.searchsorted(skey, k)
    assert np.all(skey[np.minimum(at, E - 1)] == k), 'a face walks over a pair of nodes that is not an edge'
    e = order[at]
    sign = np.where((eu[e] == a) & (ev[e] == b), 1.0, -1.0)
    return sp.csr_matrix((sign * omega[e], (face_of, e)), shape=(F, E))


def node_laplacian(B, dirichlet_idx=(), neumann_idx=(), neumann_val=()):
    """y -> dirichlet_laplacian @ y + neumann_correction with L0 = -B B^T, Dirichlet ROWS zeroed, Neumann values as a
    constant source (gds/gds.py:116-120,140,156-159).  L0 y is evaluated as -B (B^T y): the V x V product matrix of the
    reference is not formed (same value up to the order of the additions)."""
    V = B.shape[0]
    keep = np.ones(V)
    keep[np.asarray(dirichlet_idx, dtype=np.intp)] = 0.
    corr = np.zeros(V)
    corr[np.asarray(neumann_idx, dtype=np.intp)] = neumann_val
    Bt = B.T.tocsr()
    return lambda y: keep * -(B @ (Bt @ y)) + corr


def node_advect(B):
    """(v, y) -> -(B .* v) relu(-B .* sign v)^T y   (gds/gds.py:166-177; row-wise: SURVEY.md App. A.3)"""
    def f(v, y):
        v = np.asarray(v, dtype=float)
        Bp = (-(B @ sp.diags(np.sign(v)))).maximum(0)
        return -((B @ sp.diags(v)) @ (Bp.T @ y))
    return f


def edge_laplacian(B, B2, dirichlet_idx=()):
    """Hodge-1 Laplacian of gds/gds.py:296-317 without `free` / `normal`: Z_D(-B^T B y) + Z_D(-B2^T B2 y)"""
    D = np.asarray(dirichlet_idx, dtype=np.intp)
    Bt, B2t = B.T.tocsr(), B2.T.tocsr()

    def f(y):
        dd = -(Bt @ (B @ y))
        dd[D] = 0
        d_d = -(B2t @ (B2 @ y))
        d_d[D] = 0
        return dd + d_d
    return f


def step_fixed(rhs, y0, t0, h, nsteps, dirichlet_idx=(), dirichlet_at=None, scheme='rk4'):
    """`nsteps` fixed steps of one first-order state vector following fds.step_dydt / fds.dydt / apply_constraints
    (gds/fds.py:284-305,359-363; coupled: 471-507): the Dirichlet rows of every stage derivative are zero, and after each
    step they are overwritten with the boundary values at the new time (`dirichlet_at(t)` -> values; static when it
    ignores t).  `rhs(t, y_stage, y_accepted)`: operators the law calls WITHOUT an argument read `obs.y`, which is the
    solver's accepted state and only changes at the end of a step (fds.py:377-378)."""
    D = np.asarray(dirichlet_idx, dtype=np.intp)
    it = None

    def fun(t, y):
        d = np.array(rhs(t, y, it.y), dtype=float, copy=True)
        d[D] = 0.                                                      # fds.py:303-304
        return d
    cls = RK4 if scheme == 'rk4' else Euler
    it = cls(fun, t0, np.array(y0, dtype=float, copy=True), np.inf, max_step=h)
    for _ in range(nsteps):
        it.t_bound = it.t + h
        it.status = 'running'
        it.step()
        if D.size:
            it.y[D] = dirichlet_at(it.t)                               # fds.py:289-290,362
    return it.y
