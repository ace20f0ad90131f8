"""oracle/array_oracle.py (the array-form oracle used for full-size parity on the GPU) against oracle/gds_oracle.py (pinned
against the unmodified reference's golden vectors): same operators, same stepping semantics, on small graphs."""
import networkx as nx
import numpy as np
import pytest

import cases
from adaptors import rel_l2
from oracle import array_oracle as ao
from oracle import gds_oracle as go


def _arrays(low):
    faces = list(low.faces.keys())
    fptr = np.concatenate([[0], np.cumsum([len(f) for f in faces])]).astype(np.int64)
    fnodes = np.array([low.nodes[n] for f in faces for n in f], dtype=np.int64)
    return low.V, low.edge_u, low.edge_v, low.w, fptr, fnodes


@pytest.mark.parametrize('gname', ['grid_6x7', 'hex_3x4', 'hex_4x5_w', 'tri_4x6', 'tri_5x7_w'])
def test_array_operators_equal_the_object_oracle(gname):
    G = cases.OP_GRAPHS[gname]()
    if hasattr(G, 'faces'):
        del G.faces
    nd, ed = go.node_gds(G), go.edge_gds(G)
    V, eu, ev, w, fptr, fnodes = _arrays(ed.low)
    B = ao.incidence(V, eu, ev, w)
    B2 = ao.curl_matrix(V, eu, ev, w, fptr, fnodes)
    assert abs(B - ed.incidence).max() == 0
    assert abs(B2 - ed.curl_face).max() == 0
    rng = np.random.default_rng(7)
    yv, ye = rng.standard_normal(V), rng.standard_normal(len(eu))
    D_n, N_n = np.arange(0, V, 7), np.arange(3, V, 11)
    N_n = np.setdiff1d(N_n, D_n)
    nodes = list(nd.X)
    nd.set_evolution(dydt=lambda t, y: nd.laplacian(y), max_step=1e-2)
    nd.set_constraints(dirichlet={nodes[i]: 0.25 * i for i in D_n}, neumann={nodes[i]: -0.1 * (i % 5 + 1) for i in N_n})
    lap = ao.node_laplacian(B, D_n, N_n, [-0.1 * (i % 5 + 1) for i in N_n])
    assert rel_l2(lap(yv), nd.laplacian(yv)) <= 1e-14
    assert rel_l2(ao.node_advect(B)(ye, yv), nd.advect(ye, yv)) <= 1e-14
    edges = list(ed.X)
    D_e = np.arange(0, len(eu), 9)
    ed.set_evolution(dydt=lambda t, y: ed.laplacian(y), max_step=1e-2)
    ed.set_constraints(dirichlet={edges[i]: 0.1 * i for i in D_e})
    assert rel_l2(ao.edge_laplacian(B, B2, D_e)(ye), ed.laplacian(ye)) <= 1e-14


def test_array_stepping_equals_the_object_oracle():
    """the heat law of BASELINE config 1 (static + time-varying Dirichlet) stepped by both forms"""
    n, h, nsteps = 12, 1e-2, 25
    want = cases.case_heat_c1(__import__('adaptors').oracle_api(), n=n, nsteps=nsteps, every=nsteps)['y_T'][-1]
    G = nx.grid_2d_graph(n, n)
    low = go.Lowered(G, need_faces=False)
    B = ao.incidence(low.V, low.edge_u, low.edge_v, low.w)
    jj = np.arange(n)
    D = np.concatenate([jj, (n - 1) * n + jj])
    i, j = np.divmod(np.arange(n * n), n)
    y0 = np.exp(-((i - n / 2) ** 2 + (j - n / 2) ** 2) / (n / 2))
    g = lambda t: np.concatenate([np.sin(t + jj / 4) ** 2, np.zeros(n)])
    y0[D] = g(0.)
    lap = ao.node_laplacian(B, D)
    got = ao.step_fixed(lambda t, y, ya: lap(y), y0, 0., h, nsteps, D, g)
    assert rel_l2(got, want) <= 1e-13


def test_array_coupled_stepping_keeps_the_accepted_state_semantics():
    """BASELINE config 2 at small size: `c.advect(u)` (no y argument) reads the ACCEPTED state of both fields at every
    stage (fds.py:377-378), the Laplacians the stage state -- the array form against the object oracle"""
    import math
    m, n, nsteps, h, nu, kappa = 6, 8, 20, 1e-3, 0.1, 0.05
    want = cases.case_hex_advdiff_c2(__import__('adaptors').oracle_api(), m=m, n=n, nsteps=nsteps, every=nsteps)
    G = cases.g_hex(m, n)
    low = go.Lowered(G, need_faces=True)
    V, eu, ev, w, fptr, fnodes = _arrays(low)
    E = len(eu)
    B, B2 = ao.incidence(V, eu, ev, w), ao.curl_matrix(V, eu, ev, w, fptr, fnodes)
    pos = np.array([G.nodes[v]['pos'] for v in G.nodes()])
    hh = math.sqrt(3) / 2
    lo, hi = pos[:, 1].min(), pos[:, 1].max()
    bnd = (pos[:, 1] < lo + 0.6 * hh) | (pos[:, 1] > hi - 0.6 * hh)
    D_e = np.nonzero(bnd[eu] & bnd[ev])[0]
    N_n = np.nonzero(pos[:, 1] > hi - 1e-9)[0]
    elap, nlap, adv = ao.edge_laplacian(B, B2, D_e), ao.node_laplacian(B, (), N_n, -0.1), ao.node_advect(B)
    u0 = 0.1 * np.random.default_rng(0).standard_normal(E)
    c0 = np.random.default_rng(1).random(V)
    u0[D_e] = 0.

    def rhs(t, y, ya):
        return np.concatenate([nu * elap(y[:E]), -adv(ya[:E], ya[E:]) + kappa * nlap(y[E:])])
    got = ao.step_fixed(rhs, np.concatenate([u0, c0]), 0., h, nsteps, D_e, lambda t: np.zeros(D_e.size))
    assert rel_l2(got[:E], want['y_u'][-1]) <= 1e-13 and rel_l2(got[E:], want['y_c'][-1]) <= 1e-13
