"""The committed golden vectors ARE the outputs of the unmodified reference: where the reference is present
(/root/reference in the build container; its untouched copy oracle/_ref/ made by build() elsewhere) a sample of them is
regenerated in-process through tests/golden/make_golden.py's adaptor and compared with the committed files.  Skipped
when the reference is absent."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')

from oracle.ref_harness import reference_root  # noqa: E402

pytestmark = pytest.mark.skipif(reference_root() is None, reason='reference not present (/root/reference or oracle/_ref)')


@pytest.fixture(scope='module')
def ref_api():
    from oracle.ref_api import make_ref_api      # the adaptor tests/golden/make_golden.py generates the files with
    return make_ref_api()


def _same(out, path):
    gold = np.load(path)
    assert set(out) == set(gold.files)
    for k in gold.files:
        a, b = np.asarray(out[k]), gold[k]
        assert a.shape == b.shape, k
        if a.dtype.kind in 'fc':
            assert np.allclose(a, b, rtol=1e-13, atol=1e-15), k     # same code, same machine class: last-bit BLAS noise only
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize('gname', ['grid_6x7', 'hex_3x4'])
def test_operator_golden_is_fresh(ref_api, gname):
    _same(cases.operator_outputs(ref_api, gname), os.path.join(GOLD, f'ops_{gname}.npz'))


@pytest.mark.parametrize('cname', ['order2_dirichlet', 'map_time_varying_dirichlet', 'face_diffusion'])
def test_quirk_golden_is_fresh(ref_api, cname):
    _same(cases.QUIRK_CASES[cname](ref_api), os.path.join(GOLD, f'quirk_{cname}.npz'))


def test_trajectory_golden_is_fresh(ref_api):
    _same(cases.TRAJ_CASES['heat_c1_euler_small'](ref_api), os.path.join(GOLD, 'traj_heat_c1_euler_small.npz'))


@pytest.mark.parametrize('cname', ['poisson', 'stokes_hex_w'])
def test_lhs_golden_is_fresh(ref_api, cname):
    """the reference's cvx path (fds.py:97-113,307-328) with oracle/cvx_shim.py as its `cvxpy`"""
    gold = np.load(os.path.join(GOLD, f'lhs_{cname}.npz'))
    out = cases.LHS_CASES[cname](ref_api)
    assert set(out) == set(gold.files)
    for k in gold.files:       # a dense least-squares solve sits inside: LAPACK rounding, not last-bit identity
        assert np.allclose(np.asarray(out[k]), gold[k], rtol=1e-10, atol=1e-13), k


def test_cvx_stand_in_solves_the_constrained_least_squares_problem_exactly():
    """oracle/cvx_shim.py on its own: argmin |A y - f|^2 subject to y[idx] == v equals the KKT solution"""
    from oracle import cvx_shim as cp
    rng = np.random.default_rng(3)
    A, f = rng.standard_normal((9, 6)), rng.standard_normal(9)
    idx, v = np.array([1, 4]), np.array([0.3, -1.2])
    y = cp.Variable(6)
    par = cp.Parameter(2)
    par.value = v
    expr = A @ y - f
    expr2 = f - A @ y                                  # ndarray on the left
    assert np.allclose(expr2.M, -expr.M) and np.allclose(expr2.c, -expr.c)
    prob = cp.Problem(cp.Minimize(cp.sum_squares(expr)), [y[idx] == par])
    prob.solve(warm_start=True)
    assert prob.status == 'optimal'
    free = np.setdiff1d(np.arange(6), idx)
    want = np.zeros(6); want[idx] = v
    want[free] = np.linalg.solve(A[:, free].T @ A[:, free], A[:, free].T @ (f - A[:, idx] @ v))
    assert np.allclose(y.value, want, rtol=1e-12, atol=1e-13)
    # in-place arithmetic with an ndarray on the left turns the name into an expression (examples/fluid.py:28-30)
    lhs = np.ones(9)
    lhs[2] = 0.
    lhs -= (A @ y) / 2.0
    assert isinstance(lhs, cp.Affine) and np.allclose(lhs.M, -A / 2.0) and lhs.c[2] == 0. and lhs.c[0] == 1.


def test_jacobian_and_ragged_golden_are_fresh(ref_api):
    _same(cases.jacobian_outputs(ref_api, 'hex_3x4'), os.path.join(GOLD, 'jac_hex_3x4.npz'))
    _same(cases.run_ops(ref_api, cases.RAGGED['star_70_weighted']()), os.path.join(GOLD, 'ragged_star_70_weighted.npz'))


@pytest.mark.parametrize('gname', ['grid_6x7', 'hex_3x4', 'tri_4x6', 'rgg_60'])
def test_public_index_maps_and_boundary_bookkeeping_equal_the_reference(gname, monkeypatch):
    """the public attributes examples and renderers read -- X, iX, nodes, edges, faces, ndim, X_dirichlet,
    dirichlet_indices / values, X_neumann, dynamic flags (gds.py:16-55, fds.py:175-243) -- from the host mirror on a
    host-only context against the unmodified reference, same graph, same boundary conditions"""
    import math
    import gds_b200
    from oracle.ref_harness import load_reference
    ref = load_reference()
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    try:
        for kind in ('node_gds', 'edge_gds', 'face_gds'):
            a, b = getattr(gds_b200, kind)(cases.OP_GRAPHS[gname]()), getattr(ref, kind)(cases.OP_GRAPHS[gname]())
            assert a.ndim == b.ndim
            assert list(a.X.items()) == list(b.X.items())
            assert list(a.iX.items()) == list(b.iX.items())
            assert list(a.nodes.items()) == list(b.nodes.items())
            assert list(a.edges.items()) == list(b.edges.items())
            assert list(a.faces.items()) == list(b.faces.items())
            if a.ndim == 0:
                continue
            pts = list(a.X.keys())
            static = {pts[i]: 0.25 * i for i in range(0, len(pts), 5)}
            dyn = lambda t, x, pts=pts: math.sin(t + 0.1 * len(str(x))) if x in pts[2::7] else None   # noqa: E731
            for f, mod in ((a, gds_b200), (b, ref)):
                f.set_evolution(dydt=lambda t, y: 0 * y, max_step=1e-2,
                                integrator=(gds_b200.Integrators.rk4 if mod is gds_b200 else ref.Integrators.implicit_midpoint))
                f.set_constraints(dirichlet=mod.combine_bcs(static, dyn),
                                  neumann=({pts[1]: -0.5} if kind == 'node_gds' else {}))
            assert list(a.X_dirichlet) == list(b.X_dirichlet)
            assert np.array_equal(a.dirichlet_indices, b.dirichlet_indices)
            assert np.allclose(a.dirichlet_values, b.dirichlet_values, rtol=0, atol=0)
            assert list(a.X_neumann) == list(b.X_neumann) and np.array_equal(a.neumann_indices, b.neumann_indices)
            assert a.dynamic_dirichlet == b.dynamic_dirichlet and a.dynamic_neumann == b.dynamic_neumann
    finally:
        gds_b200.Context._default.clear()


@pytest.mark.parametrize('gname', list(cases.fuzz_graphs()))
def test_oracle_equals_the_live_reference_on_irregular_graphs(ref_api, gname):
    """no committed vectors here: the oracle and the unmodified reference run side by side on random / irregular graphs
    (components, isolated nodes, leaves, hubs, weights, shuffled labels), every operator and a short nonlinear trajectory"""
    from adaptors import oracle_api, rel_l2
    want = cases.run_ops(ref_api, cases.fuzz_graphs()[gname])
    got = cases.run_ops(oracle_api(), cases.fuzz_graphs()[gname])
    assert set(want) == set(got)
    for k in want:
        assert np.shape(got[k]) == np.shape(want[k]), k
        assert rel_l2(got[k], want[k]) <= 1e-12, (gname, k, rel_l2(got[k], want[k]))


def _graph_equal(A, B):
    assert list(A.nodes()) == list(B.nodes())
    assert list(A.edges()) == list(B.edges())
    pa, pb = nx_attr(A, 'pos'), nx_attr(B, 'pos')
    assert set(pa) == set(pb)
    for v in pa:
        assert np.array_equal(np.asarray(pa[v], dtype=float), np.asarray(pb[v], dtype=float)), v


def nx_attr(G, key):
    import networkx as nx
    return nx.get_node_attributes(G, key)


@pytest.mark.parametrize('maker,args,kw', [
    ('square_lattice', (5, 8), {}), ('square_lattice', (6, 4), {'diagonals': True}),
    ('triangular_lattice', (4, 7), {}), ('triangular_lattice', (5, 6), {}),
    ('hexagonal_lattice', (3, 4), {}), ('hexagonal_lattice', (2, 5), {'with_faces': False}),
])
def test_lattice_helpers_equal_the_reference(maker, args, kw, monkeypatch):
    """the domain builders of the examples (generators.py:42-200,396-429,444-534): same nodes, G.edges() order, positions,
    faces, boundary sub-graphs (with_boundaries and get_planar_boundary) and outer face as the unmodified reference"""
    import gds_b200
    from oracle.ref_harness import load_reference
    ref = load_reference()
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    try:
        Ga, sides_a = getattr(gds_b200, maker)(*args, with_boundaries=True, **kw)
        Gb, sides_b = getattr(ref, maker)(*args, with_boundaries=True, **kw)
        _graph_equal(Ga, Gb)
        for a, b in zip(sides_a, sides_b):
            assert list(a.nodes()) == list(b.nodes()) and list(a.edges()) == list(b.edges())
        assert hasattr(Ga, 'faces') == hasattr(Gb, 'faces')
        if hasattr(Gb, 'faces'):
            assert [tuple(f) for f in Ga.faces] == [tuple(f) for f in Gb.faces]
        for a, b in zip(gds_b200.get_planar_boundary(Ga), ref.get_planar_boundary(Gb)):
            assert list(a.nodes()) == list(b.nodes()) and list(a.edges()) == list(b.edges())
        fa, oa = gds_b200.embedded_faces(getattr(gds_b200, maker)(*args, **kw))
        fb, ob = ref.embedded_faces(getattr(ref, maker)(*args, **kw))
        assert [tuple(f) for f in fa] == [tuple(f) for f in fb] and tuple(oa) == tuple(ob)
        wa = gds_b200.set_edge_weights(Ga, lambda e: 1.0 + 0.1 * len(str(e)))
        wb = ref.set_edge_weights(Gb, lambda e: 1.0 + 0.1 * len(str(e)))
        assert gds_b200.get_edge_weights(wa) == ref.get_edge_weights(wb)
        assert gds_b200.get_edge_lengths(Ga) == ref.get_edge_lengths(Gb)
        assert np.array_equal(gds_b200.degree_matrix(Ga).toarray(), ref.degree_matrix(Gb).toarray())
        some = list(Ga.nodes())[::3]
        assert list(gds_b200.edge_domain(Ga, some)) == list(ref.edge_domain(Gb, some))
    finally:
        gds_b200.Context._default.clear()
