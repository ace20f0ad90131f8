"""More stepping parity against the CPU oracle: corners of the reference semantics and operator
combinations the BASELINE cases do not reach."""
import math
import os

import networkx as nx
import numpy as np
import pytest

import cases
from adaptors import oracle_api, product_api, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def both(run):
    a, b = run(product_api()), run(oracle_api())
    assert set(a) == set(b)
    for k in b:
        assert np.shape(a[k]) == np.shape(b[k]), k
        assert rel_l2(a[k], b[k]) <= TOL, (k, rel_l2(a[k], b[k]))


@pytest.mark.parametrize('cname', list(cases.QUIRK_CASES))
def test_quirk_cases(cname):
    """order-2 Dirichlet on the last block (App. B.4), recurrence maps with time-varying Dirichlet values, the smallest
    max_step of a coupled system (fds.py:437), free / normal rows and fluxes, lie_advect + bilaplacian, face fields,
    accepted-state vs stage-state operators (App. B.3): against the CPU oracle AND against golden vectors produced by the
    unmodified reference (tests/golden/quirk_*.npz)"""
    run = cases.QUIRK_CASES[cname]
    got, want = run(product_api()), run(oracle_api())
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'quirk_{cname}.npz'))
    assert set(got) == set(want) == set(gold.files)
    for k in want:
        assert np.shape(got[k]) == np.shape(want[k]), k
        assert rel_l2(got[k], want[k]) <= TOL, (k, rel_l2(got[k], want[k]))
        assert rel_l2(got[k], gold[k]) <= TOL, (k, rel_l2(got[k], gold[k]))
    if cname == 'accepted_vs_stage':
        assert rel_l2(got['accepted'], got['stage']) > 1e-3


@pytest.mark.parametrize('gname', ['hex_3x4', 'tri_5x7_w'])
def test_linearised_advection_perturbation_equation(gname):
    """dv/dt = P-free tangent dynamics J(u) v - 0.2 v around a fixed flow u (the Lyapunov analysis of
    examples/turbulence.py:252-268 as a law): advect_jvp inside a traced law, aggregates of u hoisted out of the stages"""
    def run(api):
        G = cases.OP_GRAPHS[gname]()
        v = api.edge_gds(G)
        rng = np.random.default_rng(3)
        base = rng.standard_normal(v.ndim)
        base[::5] = 0.0
        v0 = rng.standard_normal(v.ndim)
        api.evolve(v, dydt=lambda t, y: v.advect_jvp(y, base) - 0.2 * y, max_step=5e-3)
        v.set_initial(y0=lambda e: v0[v.X[e]])
        out = {}
        for i in range(3):
            v.step(0.02)
            out[f'y{i}'] = np.array(v.y, copy=True)
        return out
    both(run)


def test_navier_stokes_with_pressure_solve_every_step():
    """`lhs=` pressure field coupled to the velocity law (examples/fluid.py:14-38): device CG against the oracle's dense
    least squares.  Parity against the reference itself is unpinned for this mode (CVXPY is not installed here)."""
    a, b = cases.case_ns_pressure(product_api()), cases.case_ns_pressure(oracle_api())
    assert np.allclose(a['t'], b['t'], rtol=0, atol=1e-12)
    for i in range(b['y_velocity'].shape[0]):
        assert rel_l2(a['y_velocity'][i], b['y_velocity'][i]) <= 1e-9, rel_l2(a['y_velocity'][i], b['y_velocity'][i])
        assert rel_l2(a['y_pressure'][i], b['y_pressure'][i]) <= 1e-8, rel_l2(a['y_pressure'][i], b['y_pressure'][i])
