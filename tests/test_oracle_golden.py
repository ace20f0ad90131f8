"""The CPU oracle (oracle/gds_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Index maps must be bit-exact; float64 results within 1e-12 relative L2
(north_star bar for the product is 1e-10)."""
import os

import numpy as np
import pytest

import cases
from adaptors import oracle_api, rel_l2

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 1e-12
INT_KEYS = ('map_edges', 'map_faces_ptr', 'map_faces_nodes', 'n_faces')


@pytest.mark.parametrize('gname', list(cases.OP_GRAPHS))
def test_operators_match_reference(gname):
    gold = np.load(os.path.join(GOLD, f'ops_{gname}.npz'))
    out = cases.operator_outputs(oracle_api(), gname)
    assert set(out) == set(gold.files)
    for k in gold.files:
        if k in INT_KEYS or k == 'map_nodes_repr':
            assert np.array_equal(out[k], gold[k]), k      # index maps: bit-exact
        elif k.startswith('in_') or k == 'edge_weights':
            assert np.array_equal(out[k], gold[k]), k      # same seeded inputs
        else:
            assert rel_l2(out[k], gold[k]) <= TOL, (k, rel_l2(out[k], gold[k]))


@pytest.mark.parametrize('gname', list(cases.JAC_GRAPHS))
def test_advection_jacobian_matches_reference(gname):
    """gds.py:391-395.  The oracle rounds the operator entries to float32 exactly as the reference's torch copies do."""
    gold = np.load(os.path.join(GOLD, f'jac_{gname}.npz'))
    out = cases.jacobian_outputs(oracle_api(), gname)
    assert np.array_equal(out['in_y'], gold['in_y'])
    assert rel_l2(out['jac'], gold['jac']) <= TOL
    # and the full-precision Jacobian (what the device path computes) is the derivative of the oracle's advect
    G = cases.OP_GRAPHS[gname]()
    ed = oracle_api().edge_gds(G)
    y, live = gold['in_y'], gold['in_y'] != 0
    v = np.random.default_rng(6).standard_normal(y.size) * live
    fd = (ed.advect(y + 1e-7 * v) - ed.advect(y - 1e-7 * v)) / 2e-7
    assert rel_l2(ed.advect_jvp(v, y), fd) <= 1e-7


@pytest.mark.parametrize('cname', list(cases.QUIRK_CASES))
def test_quirk_cases_match_reference(cname):
    """corners of the reference semantics: order-2 Dirichlet on the last block, recurrence maps with
    time-varying Dirichlet values, smallest max_step of a coupled system, free / normal rows, lie_advect + bilaplacian,
    face fields, accepted-state vs stage-state operators -- each run on the UNMODIFIED reference"""
    gold = np.load(os.path.join(GOLD, f'quirk_{cname}.npz'))
    out = cases.QUIRK_CASES[cname](oracle_api())
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.shape(out[k]) == gold[k].shape, k
        assert rel_l2(out[k], gold[k]) <= TOL, (cname, k, rel_l2(out[k], gold[k]))


@pytest.mark.parametrize('gname', list(cases.RAGGED))
def test_ragged_and_tiny_graphs_match_reference(gname):
    """a single node, isolated nodes, one edge, paths and cycles off the 32 / 128-row boundaries, stars (degree-1 nodes:
    the inf -> 0 rule of the dual weights, gds.py:253-257), a complete graph, a grid with isolated nodes -- every operator
    and a few RK4 steps of a nonlinear law, each run on the UNMODIFIED reference"""
    gold = np.load(os.path.join(GOLD, f'ragged_{gname}.npz'))
    out = cases.run_ops(oracle_api(), cases.RAGGED[gname]())
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.shape(out[k]) == gold[k].shape, k
        assert rel_l2(out[k], gold[k]) <= TOL, (gname, k, rel_l2(out[k], gold[k]))


@pytest.mark.parametrize('cname', list(cases.API_CASES))
def test_api_surface_matches_reference(cname):
    """System.solve, project=, fields outside the system, oriented point reads, curl of an evolved field, single-edge
    differences -- each run on the UNMODIFIED reference"""
    gold = np.load(os.path.join(GOLD, f'api_{cname}.npz'))
    out = cases.API_CASES[cname](oracle_api())
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.shape(out[k]) == gold[k].shape, k
        assert rel_l2(out[k], gold[k]) <= TOL, (cname, k, rel_l2(out[k], gold[k]))


@pytest.mark.parametrize('cname', list(cases.TRAJ_CASES))
def test_trajectories_match_reference(cname):
    gold = np.load(os.path.join(GOLD, f'traj_{cname}.npz'))
    out = cases.TRAJ_CASES[cname](oracle_api())
    assert set(out) == set(gold.files)
    assert np.allclose(out['t'], gold['t'], rtol=0, atol=1e-12)
    for k in gold.files:
        if k == 't':
            continue
        assert np.all(np.isfinite(gold[k]))
        for i in range(gold[k].shape[0]):
            err = rel_l2(out[k][i], gold[k][i])
            assert err <= TOL, (cname, k, i, err)


@pytest.mark.parametrize('cname', list(cases.LHS_CASES))
def test_lhs_fields_match_the_reference_cvx_path(cname):
    """`set_evolution(lhs=...)` (fds.py:97-113,196-202,307-328; the pressure field of examples/fluid.py:14-38), alone and
    coupled to a velocity field, static and time-varying Dirichlet values: golden vectors from the UNMODIFIED reference
    running its own cvx path, with oracle/cvx_shim.py standing in for the absent CVXPY (the reference assembles the
    programme -- objective from its operators, Dirichlet constraint, solves before every RHS evaluation of the coupled
    system -- and the stand-in minimises it exactly)"""
    gold = np.load(os.path.join(GOLD, f'lhs_{cname}.npz'))
    out = cases.LHS_CASES[cname](oracle_api())
    assert set(out) == set(gold.files)
    assert np.allclose(out['t'], gold['t'], rtol=0, atol=1e-12)
    for k in gold.files:
        if k == 't':
            continue
        assert np.all(np.isfinite(gold[k])) and np.linalg.norm(gold[k]) > 0
        for i in range(gold[k].shape[0]):
            err = rel_l2(out[k][i], gold[k][i])
            assert err <= TOL, (cname, k, i, err)
