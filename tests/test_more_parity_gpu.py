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
    least squares (the golden vectors of the reference's own cvx path are compared in the next test)."""
    a, b = cases.case_ns_pressure(product_api()), cases.case_ns_pressure(oracle_api())
    assert np.allclose(a['t'], b['t'], rtol=0, atol=1e-12)
    for i in range(b['y_velocity'].shape[0]):
        assert rel_l2(a['y_velocity'][i], b['y_velocity'][i]) <= 1e-9, rel_l2(a['y_velocity'][i], b['y_velocity'][i])
        assert rel_l2(a['y_pressure'][i], b['y_pressure'][i]) <= 1e-8, rel_l2(a['y_pressure'][i], b['y_pressure'][i])


@pytest.mark.parametrize('cname', list(cases.LHS_CASES))
def test_lhs_fields_match_the_reference_golden_vectors(cname):
    """the device solve of `lhs=` fields against golden vectors of the UNMODIFIED reference's cvx path (its CVXPY replaced
    by the exact least-squares stand-in oracle/cvx_shim.py, tests/golden/make_golden.py): the coupled Navier-Stokes case of
    the test above, a Poisson field stepped on its own with time-varying Dirichlet values, Stokes flow on a weighted
    hexagonal lattice with a time-varying pressure drop"""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', f'lhs_{cname}.npz'))
    out = cases.LHS_CASES[cname](product_api())
    assert np.allclose(out['t'], gold['t'], rtol=0, atol=1e-12)
    for k in gold.files:
        if k == 't':
            continue
        tol = 2e-8 if 'pressure' in k or k == 'y_p' else 2e-9
        for i in range(gold[k].shape[0]):
            assert rel_l2(out[k][i], gold[k][i]) <= tol, (cname, k, i, rel_l2(out[k][i], gold[k][i]))


COLLAPSIBLE = {
    # laws that read accepted state only (operators without an argument): the idiom of examples/fluid.py:33-34, wave.py:14
    'wave_order2': lambda api: cases.case_wave_c5(api, dims=(5, 6, 9), nsteps=12, every=4),
    'ns_cavity': lambda api: cases.case_ns_cavity_c3(api, m=6, n=10, nsteps=12, every=4),
    'order2_dirichlet': lambda api: cases.QUIRK_CASES['order2_dirichlet'](api),
}


@pytest.mark.parametrize('cname', list(COLLAPSIBLE))
def test_collapsed_rk4_is_bit_identical_to_the_staged_form(cname, monkeypatch):
    """a law without stage-state operands has k1 = k2 = k3 = k4: the engine runs its passes once per step and combines
    the four stages in registers (gds_system_set_collapsed) -- same operations in the same order as the four stage
    passes, so the results must agree to the last bit (and with the oracle to 1e-10)"""
    from gds_b200 import engine as E
    seen = []
    build = E.StepEngine._build

    def spy(self):
        build(self)
        seen.append((self.collapsed, self.n_stages, self.kernels_per_step))
    monkeypatch.setattr(E.StepEngine, '_build', spy)
    fast = COLLAPSIBLE[cname](product_api())
    assert seen and all(c and n == 1 for c, n, _ in seen), seen
    k_fast = seen[-1][2]
    del seen[:]
    monkeypatch.setenv('GDS_B200_NO_COLLAPSE', '1')
    slow = COLLAPSIBLE[cname](product_api())
    assert seen and all((not c) and n == 4 for c, n, _ in seen), seen
    assert k_fast < seen[-1][2]
    want = COLLAPSIBLE[cname](oracle_api())
    for k in want:
        assert np.array_equal(np.asarray(fast[k]), np.asarray(slow[k])), (cname, k)
        assert rel_l2(fast[k], want[k]) <= TOL, (cname, k)


def test_laws_that_read_the_stage_state_are_not_collapsed():
    T = __import__('gds_b200').node_gds(cases.g_grid(8, 8))
    T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
    T.set_initial(y0=lambda x: float(x[0]))
    assert not T._ensure_engine().collapsed
    S = __import__('gds_b200').node_gds(cases.g_grid(8, 8))
    S.set_evolution(dydt=lambda t, y: S.laplacian() * np.cos(t), max_step=1e-2)      # stage time
    S.set_initial(y0=lambda x: float(x[0]))
    assert not S._ensure_engine().collapsed


def test_coupled_reset_goes_back_to_the_state_taken_at_couple_time():
    """coupled_fds.reset (fds.py:509-527): the shared solver restarts at t0 from the state concatenated at couple() time,
    recurrence maps are reset one by one"""
    def run(api):
        G = cases.g_grid(6, 5)
        T, m = api.node_gds(G), api.node_gds(G)
        api.evolve(T, dydt=lambda t, y: T.laplacian(y) + 0.5 * m.y, max_step=1e-2)
        api.evolve_map(m, map_fun=lambda t, y: 0.9 * y + 0.2 * T.y, dt=0.025)
        T.set_initial(y0=lambda x: math.sin(x[0]) + 0.1 * x[1])
        m.set_initial(y0=lambda x: 0.05 * x[0] - 0.02 * x[1])
        T.set_constraints(dirichlet=lambda x: 1.0 if x[0] == 0 else None)
        sysobj = api.couple({'T': T, 'm': m})
        out = {}
        for rnd in range(2):
            for i in range(4):
                sysobj.step(0.02)
            out[f'T_{rnd}'], out[f'm_{rnd}'], out[f't_{rnd}'] = np.array(T.y), np.array(m.y), np.array([sysobj.t])
            sysobj.stepper.reset()
            out[f'T_reset_{rnd}'], out[f't_reset_{rnd}'] = np.array(T.y), np.array([sysobj.t])
        return out
    both(run)
    a = run(product_api())
    assert np.array_equal(a['T_reset_0'], a['T_reset_1']) and a['t_reset_0'][0] == 0.


def _cml_with_dirichlet(api):
    """coupled map lattice y <- f(y) + (eps/8) L f(y), f(y) = 1 - 1.7 y^2, with a time-varying Dirichlet row set and a host
    write in the middle (the carried operand f(y) must follow both)"""
    G = cases.g_rgg(700, 0.08, 5)
    x = api.node_gds(G)
    api.evolve_map(x, map_fun=lambda t, y: (1 - 1.7 * y ** 2) + (0.3 / 8.0) * x.laplacian(1 - 1.7 * y ** 2), dt=1.0)
    y0 = np.random.default_rng(2).uniform(-1, 1, 700)
    x.set_initial(y0=lambda v: y0[v])
    x.set_constraints(dirichlet=lambda t, v: 0.5 * math.sin(0.7 * t) if v % 13 == 0 else None)
    out = {}
    for i in range(6):
        x.step(1.0)
        out[f'y{i}'] = np.array(x.y, copy=True)
    return out


def test_carried_gather_operand_is_bit_identical_to_on_the_fly_evaluation(monkeypatch):
    """maps of the form y <- ... L f(y) ...: the engine keeps Z = f(Y) beside the state (gds_system_set_shadow) and gathers
    that, instead of evaluating f for every gathered neighbour -- same f, evaluated once per row: identical bits"""
    from gds_b200 import engine as E
    seen = []
    build = E.StepEngine._build

    def spy(self):
        build(self)
        seen.append(self.modes['shadow'])
    monkeypatch.setattr(E.StepEngine, '_build', spy)
    fast = _cml_with_dirichlet(product_api())
    fast2 = cases.case_cml_c4(product_api(), n=3000, iters=8, every=4)
    assert seen == [True, True], seen
    monkeypatch.setenv('GDS_B200_NO_SHADOW', '1')
    slow = _cml_with_dirichlet(product_api())
    slow2 = cases.case_cml_c4(product_api(), n=3000, iters=8, every=4)
    assert seen[2:] == [False, False], seen
    want = _cml_with_dirichlet(oracle_api())
    for k in want:
        assert np.array_equal(fast[k], slow[k]), k
        assert rel_l2(fast[k], want[k]) <= TOL, k
    assert np.array_equal(fast2['y_x'], slow2['y_x'])


def test_phased_edge_term_kernel_is_bit_identical_to_the_term_by_term_kernel(monkeypatch):
    """edge_terms_kernel issues the loads of all terms level by level before consuming them; same arithmetic per term,
    same order of accumulation as terms_kernel: identical bits (Navier-Stokes edge law: advect + grad p + Hodge Laplacian;
    Hodge Laplacian with Dirichlet rows; on unit-weight and weighted graphs)"""
    import gds_b200

    def run():
        out = dict(cases.case_ns_cavity_c3(product_api(), m=7, n=10, nsteps=6, every=3))
        out.update({'hx_' + k: v for k, v in cases.case_hex_advdiff_c2(product_api(), m=6, n=8, nsteps=6, every=3).items()})
        G = cases.add_weights(cases.g_tri(5, 7), 12)
        u = gds_b200.edge_gds(G)
        y = np.random.default_rng(5).standard_normal(u.ndim)
        y[::7] = 0.0
        out['w_lap'], out['w_adv'], out['w_dd'] = u.laplacian(y), u.advect(y), u.dd_(y)
        return out
    fast = run()
    monkeypatch.setenv('GDS_B200_NO_PHASED', '1')
    slow = run()
    for k in fast:
        assert np.array_equal(np.asarray(fast[k]), np.asarray(slow[k])), k
