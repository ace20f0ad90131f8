"""Operator assembly on the host (C++ lowering behind the C-ABI, host-only context) against the oracle, whose index
maps are pinned to the unmodified reference by tests/golden: node / edge / face maps and the B1 / B2 incidence
structure must be bit-exact.  Also the host-side BC bookkeeping and initial state against the oracle."""
import math
import os

import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from adaptors import oracle_api, product_api


@pytest.fixture(autouse=True)
def host_only(monkeypatch):
    import gds_b200
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    yield
    gds_b200.Context._default.clear()


def csr_from_export(ptr, idx, sign, shape):
    return sp.csr_matrix((sign.astype(float), idx, ptr), shape=shape)


@pytest.mark.parametrize('gname', list(cases.OP_GRAPHS))
def test_index_maps_and_incidence_bit_exact(gname):
    import gds_b200
    from oracle import gds_oracle
    G = cases.OP_GRAPHS[gname]()
    low = gds_b200.LoweredGraph.of(G, need_faces=True)
    ref = gds_oracle.Lowered(cases.OP_GRAPHS[gname](), need_faces=True)
    assert list(low.nodes.items()) == list(ref.nodes.items())
    assert list(low.edges.items()) == list(ref.edges.items())
    assert list(low.faces.items()) == list(ref.faces.items())
    assert np.array_equal(low.weights, ref.w)
    V, E, F = ref.V, ref.E, ref.F
    # B1 rows: structure and orientation (values are sign * sqrt(w) on both sides)
    p, i, s = low.dev.export_csr(0)
    B = csr_from_export(p, i, s, (V, E))
    Bref = ref.B.copy()
    Bref.data = np.sign(Bref.data)
    assert (B != Bref).nnz == 0
    if F:
        p, i, s = low.dev.export_csr(1)
        C = csr_from_export(p, i, s, (F, E))
        Cref = ref.curl_matrix().tocsr()
        Cref.data = np.sign(Cref.data)
        assert (C != Cref).nnz == 0
        p, i, s = low.dev.export_csr(2)
        Ct = csr_from_export(p, i, s, (E, F))
        assert (Ct != Cref.T.tocsr()).nnz == 0
        # curl o grad = 0 on the lattices
        assert abs(C @ B.T).sum() == 0


@pytest.mark.parametrize('gname', list(cases.OP_GRAPHS))
def test_index_maps_match_golden(gname):
    import gds_b200
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'ops_{gname}.npz'))
    G = cases.OP_GRAPHS[gname]()
    ed = gds_b200.edge_gds(G)
    nd = gds_b200.node_gds(G)
    edges = np.array([[nd.X[a], nd.X[b]] for (a, b) in ed.X.keys()], dtype=np.int64).reshape(ed.ndim, 2)
    assert np.array_equal(edges, gold['map_edges'])
    assert np.array_equal(np.array([repr(n) for n in nd.X.keys()]), gold['map_nodes_repr'])
    assert np.array_equal(np.asarray(ed.edge_weights), gold['edge_weights'])
    assert len(ed.faces) == int(gold['n_faces'][0])
    if len(ed.faces):
        faces = list(ed.faces.keys())
        assert np.array_equal(np.cumsum([0] + [len(f) for f in faces]), gold['map_faces_ptr'])
        assert np.array_equal(np.array([nd.X[n] for f in faces for n in f]), gold['map_faces_nodes'])


def _setup_heat(api, n=12):
    G = cases.g_grid(n, n)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
    T.set_initial(y0=lambda x: math.exp(-((x[0] - 6) ** 2 + (x[1] - 6) ** 2) / 6))
    T.set_constraints(dirichlet=api.combine_bcs(lambda x: 0. if x[0] == n - 1 else None,
                                                lambda t, x: math.sin(t + x[1] / 4) ** 2 if x[0] == 0 else None),
                      neumann=lambda x: -0.1 if (x[1] == 0 and 0 < x[0] < n - 1) else None)
    return T


def test_bc_bookkeeping_and_initial_state_match_oracle():
    a, b = _setup_heat(product_api()), _setup_heat(oracle_api())
    assert a.dynamic_dirichlet == b.dynamic_dirichlet and a.dynamic_neumann == b.dynamic_neumann
    assert np.array_equal(a.dirichlet_indices, b.dirichlet_indices)
    assert np.array_equal(a.dirichlet_values, b.dirichlet_values)
    assert np.array_equal(a.neumann_indices, b.neumann_indices)
    assert np.array_equal(a.neumann_values, b.neumann_values)
    assert list(a.X_dirichlet) == list(b.X_dirichlet)
    assert np.array_equal(a.neumann_correction, b.neumann_correction)
    assert np.array_equal(a.y, b.y)          # host state before the first step
    assert a.t == b.t == 0.


def test_order2_dirichlet_lands_on_last_block_like_the_reference():
    def setup(api):
        G = cases.g_grid(5, 4)
        T = api.node_gds(G)
        api.evolve(T, dydt=lambda t, y: T.laplacian(), order=2, max_step=1e-2)
        T.set_initial(y0=lambda x: float(x[0]))
        T.set_constraints(dirichlet={(0, 0): 7.0, (2, 3): -1.0})
        return T
    a, b = setup(product_api()), setup(oracle_api())
    assert np.array_equal(a._host_state, b.integrator.y)


def test_overlapping_conditions_are_rejected():
    for api in (product_api(), oracle_api()):
        T = api.node_gds(cases.g_grid(4, 4))
        api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
        with pytest.raises(AssertionError):
            T.set_constraints(dirichlet={(0, 0): 1.0}, neumann={(0, 0): 1.0})


def test_api_preconditions():
    import gds_b200
    T = gds_b200.node_gds(cases.g_grid(3, 3))
    with pytest.raises(AssertionError):
        T.set_initial(y0=lambda x: 0.)          # fds.py:149
    with pytest.raises(AssertionError):
        T.set_constraints(dirichlet={})         # fds.py:189
    with pytest.raises(AssertionError):
        T.set_evolution()                       # fds.py:72
    with pytest.raises(ValueError):
        T.set_evolution(dydt=lambda t, y: y, integrator=gds_b200.Integrators.lsoda)   # fds.py:95
    with pytest.raises(Exception):
        T.step(0.1)                             # no evolution law


def test_internal_renumbering_keeps_public_maps_and_row_contents():
    """graphs without index locality (random geometric graphs) are renumbered internally along a Z-order curve; the
    public index maps do not change and every device row holds the same (edge, sign) list as its public row"""
    import gds_b200
    from oracle import gds_oracle
    G = cases.g_rgg(500, 0.08, 9)
    low = gds_b200.LoweredGraph.of(G)
    ref = gds_oracle.Lowered(cases.g_rgg(500, 0.08, 9), need_faces=False)
    assert low.i2p is not None and sorted(low.i2p) == list(range(500))
    assert list(low.nodes.items()) == list(ref.nodes.items()) and list(low.edges.items()) == list(ref.edges.items())
    p, i, s = low.dev.export_csr(0)
    Bref = ref.B.tocsr()
    for r_int in range(500):
        r_pub = low.i2p[r_int]
        cols = Bref.indices[Bref.indptr[r_pub]:Bref.indptr[r_pub + 1]]
        sg = np.sign(Bref.data[Bref.indptr[r_pub]:Bref.indptr[r_pub + 1]])
        order = np.argsort(cols)
        assert np.array_equal(i[p[r_int]:p[r_int + 1]], cols[order])
        assert np.array_equal(s[p[r_int]:p[r_int + 1]], sg[order])
    # vectors and row sets travel through the permutation and back
    x = np.arange(500.0)
    assert np.array_equal(low.from_dev(gds_b200._lib.NODES, low.to_dev(gds_b200._lib.NODES, x)), x)
    assert np.array_equal(low.i2p[low.rows_to_dev(gds_b200._lib.NODES, [3, 77, 400])], [3, 77, 400])
    # lattices keep their numbering
    assert gds_b200.LoweredGraph.of(cases.g_hex(6, 8)).i2p is None
    # locality: the renumbered edge list is far more banded than the public one
    d_pub = np.mean(np.abs(low.edge_u.astype(np.int64) - low.edge_v))
    d_int = np.mean(np.abs(low.p2i[low.edge_u] - low.p2i[low.edge_v]))
    assert d_int < 0.5 * d_pub


def test_graph_edited_after_a_field_was_built_is_lowered_again():
    import gds_b200
    G = cases.g_grid(4, 4)
    a = gds_b200.node_gds(G)
    b = gds_b200.node_gds(G)
    assert a._low is b._low                       # one lowering per graph
    cases.add_weights(G, 3)
    c = gds_b200.node_gds(G)
    assert c._low is not a._low and not np.all(c.edge_weights == 1.0)
    G.add_edge((0, 0), (3, 3))
    assert gds_b200.node_gds(G)._low.E == a._low.E + 1
