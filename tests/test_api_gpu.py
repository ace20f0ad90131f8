"""The rest of the drop-in surface on the device path: reset, System.solve / system(name), derived observables
(project), point reads, oriented edge reads, single-edge differences, vectorised boundary specs and native graphs
driven through the same field API -- each against the CPU oracle."""
import math

import networkx as nx
import numpy as np
import pytest

import cases
from adaptors import oracle_api, product_api, rel_l2

pytestmark = pytest.mark.gpu


def heat(api, n=10):
    G = cases.g_grid(n, n)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=2e-3)
    T.set_initial(y0=lambda x: math.exp(-((x[0] - 4) ** 2 + (x[1] - 5) ** 2) / 6))
    T.set_constraints(dirichlet=lambda t, x: math.cos(t) if x[0] == 0 else None)
    return T


def test_reset_returns_to_initial_conditions():
    T = heat(product_api())
    y0 = np.array(T.y, copy=True)
    T.step(0.01)
    y1 = np.array(T.y, copy=True)
    assert T.t == pytest.approx(0.01) and not np.array_equal(y0, y1)
    T.reset()
    assert T.t == 0. and np.array_equal(np.asarray(T.y), y0)
    T.step(0.01)
    assert np.array_equal(np.asarray(T.y), y1)         # deterministic replay


def test_system_solve_matches_oracle():
    def run(api):
        T = heat(api)
        return T.system('T').solve(0.02, 0.004)
    (ta, da), (tb, db) = run(product_api()), run(oracle_api())
    assert np.allclose(ta, tb, rtol=0, atol=1e-12)
    assert da['T'].shape == db['T'].shape
    for i in range(db['T'].shape[0]):
        assert rel_l2(da['T'][i], db['T'][i]) <= 1e-10


def test_point_reads_projections_and_partials():
    import gds_b200
    G = cases.add_weights(cases.g_hex(3, 4), 2)
    u, ref_u = gds_b200.edge_gds(G), oracle_api().edge_gds(cases.add_weights(cases.g_hex(3, 4), 2))
    rng = np.random.default_rng(3)
    y0 = rng.standard_normal(u.ndim)
    for f in (u, ref_u):
        f.set_evolution(dydt=lambda t, y, f=f: 0.1 * f.laplacian(y), max_step=1e-3)
        f.set_initial(y0=lambda e, f=f: y0[f.X[e]])
    u.step(3e-3); ref_u.step(3e-3)
    e = list(u.X.keys())[5]
    assert u(e) == pytest.approx(ref_u(e), rel=1e-10)
    assert u((e[1], e[0])) == pytest.approx(-ref_u(e), rel=1e-10)       # oriented read (gds.py:273-274)
    assert u[5] == pytest.approx(ref_u.y[5], rel=1e-10)
    # derived observable (types.py:95-105): vorticity = curl of the velocity field, evaluated lazily on the device
    vort = u.project(gds_b200.GraphObservable, lambda obs: obs.curl(), G, gds_b200.GraphDomain.faces)
    assert vort.t == u.t and rel_l2(vort.y, ref_u.curl()) <= 1e-10
    c, ref_c = gds_b200.node_gds(G), oracle_api().node_gds(cases.add_weights(cases.g_hex(3, 4), 2))
    x = rng.standard_normal(c.ndim)
    assert c.partial(e, x) == pytest.approx(ref_c.grad(x)[ref_u.X[e]], rel=1e-12)   # gds.py:149-150


def test_derived_observables_reduce_on_the_device(monkeypatch):
    """the projections of gds/examples/lid_driven_cavity.py:95-101 -- kinetic energy, momentum, rotational energy, angular
    momentum: `(v.y ** 2).sum()`, `np.abs(v.y).sum()`, `(v.curl() ** 2).sum()`, `np.abs(v.curl()).sum()` -- evaluated by
    device passes over the state where it lives plus a device reduction (gds_buf_reduce); only the scalar comes back"""
    import gds_b200
    from gds_b200 import device as D
    G = cases.add_weights(cases.g_tri(5, 7), 4)
    v, ref_v = gds_b200.edge_gds(G), oracle_api().edge_gds(cases.add_weights(cases.g_tri(5, 7), 4))
    y0 = np.random.default_rng(8).standard_normal(v.ndim)
    for f in (v, ref_v):
        f.set_evolution(dydt=lambda t, y, f=f: 0.05 * f.laplacian(y) - 0.1 * y, max_step=1e-3)
        f.set_initial(y0=lambda e, f=f: y0[f.X[e]])
    v.step(4e-3); ref_v.step(4e-3)
    views = {'kinetic energy': (lambda o: (v.y ** 2).sum(), (ref_v.y ** 2).sum()),
             'momentum': (lambda o: np.abs(v.y).sum(), np.abs(ref_v.y).sum()),
             'rotational energy': (lambda o: (v.curl() ** 2).sum(), (ref_v.curl() ** 2).sum()),
             'angular momentum': (lambda o: np.abs(v.curl()).sum(), np.abs(ref_v.curl()).sum()),
             'half energy': (lambda o: 0.5 * np.sum(v.y * v.y), 0.5 * np.sum(ref_v.y * ref_v.y)),
             'peak': (lambda o: np.abs(v.y).max() - v.y.min(), np.abs(ref_v.y).max() - ref_v.y.min()),
             'mean': (lambda o: v.y.mean(), ref_v.y.mean())}
    calls, downloads = [], []
    red, down = D.Buffer.reduce, D.Buffer.download
    monkeypatch.setattr(D.Buffer, 'reduce', lambda self, *a, **k: (calls.append(1), red(self, *a, **k))[1])
    monkeypatch.setattr(D.Buffer, 'download', lambda self, *a, **k: (downloads.append(self.n), down(self, *a, **k))[1])
    for name, (view, want) in views.items():
        obs = v.project(gds_b200.PointObservable, view)
        got = obs.y
        assert isinstance(got, float) and got == pytest.approx(want, rel=1e-10), name
        assert obs.t == v.t
    assert len(calls) == 8 and not downloads          # every view ended in a device reduction; no vector was copied back
    # a view that is arbitrary host code still works (on a copy of the state, as in the reference)
    norm = v.project(gds_b200.PointObservable, lambda o: float(np.linalg.norm(np.asarray(o.y))))
    assert norm.y == pytest.approx(float(np.linalg.norm(ref_v.y)), rel=1e-10)


def test_matrix_operands_equal_the_column_loop_bit_for_bit():
    """operators applied to a matrix (`laplacian(np.eye(n))`, gds/examples/fluid.py:350; blocks of tangent vectors for
    advect_jvp): one lowering, one upload, one device call for all columns (gds_passes_run_cols), one download -- the same
    kernels per column, so the result equals the column-by-column application to the last bit"""
    import gds_b200
    G = cases.add_weights(cases.g_hex(4, 5), 9)
    v, c = gds_b200.edge_gds(G), gds_b200.node_gds(G)
    rng = np.random.default_rng(12)
    Me, Mn = rng.standard_normal((v.ndim, 7)), rng.standard_normal((c.ndim, 5))
    y = rng.standard_normal(v.ndim)
    for got, one in ((v.laplacian(Me), lambda j: v.laplacian(Me[:, j])),          # three passes per column
                     (v.div(Me), lambda j: v.div(Me[:, j])),
                     (v.advect_jvp(Me, y), lambda j: v.advect_jvp(Me[:, j], y)),   # matrix and vector operand mixed
                     (c.laplacian(Mn), lambda j: c.laplacian(Mn[:, j])),
                     (c.bilaplacian(Mn), lambda j: c.bilaplacian(Mn[:, j])),
                     (c.grad(Mn), lambda j: c.grad(Mn[:, j]))):
        assert got.ndim == 2
        for j in range(got.shape[1]):
            assert np.array_equal(got[:, j], one(j))
    I = c.laplacian(np.eye(c.ndim))                      # the matrix of the operator itself
    assert np.array_equal(I, I.T) and np.allclose(I.sum(axis=0), 0.0, atol=1e-12)
    # a graph whose node numbering is permuted internally (no locality): rows come back in public order
    R = cases.g_rgg(300, 0.12, 3)
    x = gds_b200.node_gds(R)
    Mr = rng.standard_normal((300, 4))
    got = x.laplacian(Mr)
    for j in range(4):
        assert np.array_equal(got[:, j], x.laplacian(Mr[:, j]))


def test_native_graph_and_boundary_spec_equal_networkx_and_callables():
    """the same heat problem through nx.Graph + per-point callables and through NativeGraph + vectorised BoundarySpec"""
    import gds_b200
    n = 37
    T1 = gds_b200.node_gds(cases.g_grid(n, n))
    T1.set_evolution(dydt=lambda t, y: T1.laplacian(y), max_step=1e-2, integrator=gds_b200.Integrators.rk4)
    T1.set_initial(y0=lambda x: math.exp(-((x[0] - 18) ** 2 + (x[1] - 18) ** 2) / 20))
    T1.set_constraints(dirichlet=gds_b200.combine_bcs(lambda x: 0. if x[0] == n - 1 else None,
                                                      lambda t, x: math.sin(t + x[1] / 4) ** 2 if x[0] == 0 else None))
    T2 = gds_b200.node_gds(gds_b200.NativeGraph('grid_2d', n, n))
    T2.set_evolution(dydt=lambda t, y: T2.laplacian(y), max_step=1e-2, integrator=gds_b200.Integrators.rk4)
    i, j = np.divmod(np.arange(n * n), n)
    T2.set_initial(y0=np.exp(-((i - 18) ** 2 + (j - 18) ** 2) / 20))
    jj = np.arange(n)
    T2.set_constraints(dirichlet=gds_b200.BoundarySpec(
        np.concatenate([jj, (n - 1) * n + jj]), fun=lambda t: np.concatenate([np.sin(t + jj / 4) ** 2, np.zeros(n)])))
    for _ in range(5):
        T1.step(3e-2); T2.step(3e-2)
    assert T1.t == T2.t
    assert rel_l2(T1.y, T2.y) <= 1e-13                 # math.exp / math.sin vs their numpy counterparts: last-ulp inputs
    assert list(T2.X.keys())[:3] == [(0, 0), (0, 1), (0, 2)] and T2((1, 2)) == pytest.approx(T1((1, 2)), rel=1e-12)


def test_user_projection_runs_on_the_host_every_step():
    """project= is arbitrary host code (fds.py:363): supported through one host round trip per step"""
    def run(api):
        T = heat(api, n=8)
        T.set_constraints(dirichlet={(0, 0): 1.0}, project=lambda y: np.clip(y, 0.0, 0.6))
        T.step(0.01)
        return np.array(T.y, copy=True)
    assert rel_l2(run(product_api()), run(oracle_api())) <= 1e-10


def test_external_field_and_several_max_steps_per_call():
    """a law may read a field outside its system (uploaded when it changes); step(dt) spans several max_steps"""
    def run(api):
        G = cases.g_grid(7, 9)
        src, T = api.node_gds(G), api.node_gds(G)
        api.evolve_nil(src)
        src.set_initial(y0=lambda x: 0.1 * x[0] - 0.05 * x[1])
        api.evolve(T, dydt=lambda t, y: T.laplacian(y) + 2.0 * src.y - y * t, max_step=3e-3)
        T.set_initial(y0=lambda x: float(x[0] == 3))
        for _ in range(3):
            T.step(1e-2)
        return np.array(T.y, copy=True), T.t
    (a, ta), (b, tb) = run(product_api()), run(oracle_api())
    assert ta == pytest.approx(tb, abs=1e-15) and rel_l2(a, b) <= 1e-10


def test_leray_projection_matches_dense_pinv_and_is_a_projection():
    """edge_gds.leray_project (gds.py:249,414-419): device CG on B1 B1^T against the oracle's dense pinv (the golden
    files pin the same function to the unmodified reference at small size), plus divergence-freeness and idempotence"""
    import gds_b200
    for make in (lambda: cases.add_weights(cases.g_tri(9, 12), 4), lambda: cases.g_hex(7, 9), lambda: cases.g_rgg(300, 0.12, 1)):
        u, ref = gds_b200.edge_gds(make()), oracle_api().edge_gds(make())
        y = np.random.default_rng(8).standard_normal(u.ndim)
        z, info = u.leray_project(y, return_info=True)
        assert info['relative_residual'] <= 1e-12
        assert rel_l2(z, ref.leray_project(y)) <= 1e-10
        assert np.max(np.abs(u.div(z))) <= 1e-10 * np.max(np.abs(u.div(y)))
        assert rel_l2(u.leray_project(z), z) <= 1e-10
        # gradients are annihilated, divergence-free fields are kept
        c = gds_b200.node_gds(u.G)
        g = c.grad(np.random.default_rng(9).standard_normal(c.ndim))
        assert np.linalg.norm(u.leray_project(g)) <= 1e-10 * np.linalg.norm(g)
    with pytest.raises(NotImplementedError):
        u.set_evolution(dydt=lambda t, yy: u.laplacian(yy))
        u.set_constraints(dirichlet={list(u.X.keys())[0]: 1.0})
        u.leray_project(y)


def test_solve_records_on_the_device_and_matches_the_stepping_loop():
    """System.solve keeps its snapshots on the device (one read-back); same numbers as stepping and copying by hand,
    for a coupled system, a stride, and several internal steps per call"""
    import gds_b200

    def build():
        G = cases.g_hex(4, 5)
        u, c = gds_b200.edge_gds(G), gds_b200.node_gds(G)
        u.set_evolution(dydt=lambda t, y: 0.1 * u.laplacian(y), max_step=1e-3, integrator=gds_b200.Integrators.rk4)
        c.set_evolution(dydt=lambda t, y: -c.advect(u) + 0.05 * c.laplacian(y), max_step=1e-3,
                        integrator=gds_b200.Integrators.rk4)
        r = np.random.default_rng(11)
        u.set_initial(y0=0.1 * r.standard_normal(u.ndim))
        c.set_initial(y0=r.random(c.ndim))
        c.set_constraints(dirichlet=lambda t, x: math.sin(5 * t) if x == list(c.X)[0] else None)
        return gds_b200.couple({'u': u, 'c': c}), u, c
    s1, u1, c1 = build()
    time, data = s1.solve(0.0199, 2.5e-3, every=3)          # 8 calls of 3 internal steps; snapshots after calls 3, 6, 8
    assert data['u'].shape == (4, u1.ndim) and data['c'].shape == (4, c1.ndim)
    assert np.allclose(time, [0.0, 7.5e-3, 15e-3, 20e-3], rtol=0, atol=1e-12)
    s2, u2, c2 = build()
    want_u, want_c = [np.array(u2.y, copy=True)], [np.array(c2.y, copy=True)]
    for i in range(1, 9):
        s2.step(2.5e-3)
        if i % 3 == 0 or i == 8:
            want_u.append(np.array(u2.y, copy=True)); want_c.append(np.array(c2.y, copy=True))
    assert np.array_equal(data['u'], np.array(want_u)) and np.array_equal(data['c'], np.array(want_c))
    assert s1.t == pytest.approx(s2.t, abs=1e-15)
    assert np.array_equal(np.asarray(u1.y), np.asarray(u2.y))          # the fields carry on from the end of solve()
    s1.step(2.5e-3); s2.step(2.5e-3)
    assert np.array_equal(np.asarray(c1.y), np.asarray(c2.y))


def test_solve_to_disk_layout(tmp_path):
    """system.py:59-95: one array [n_steps, ndim] per observable with the states after every step, and 't'"""
    def make(api):
        G = cases.g_grid(6, 5)
        T = api.node_gds(G)
        api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
        T.set_initial(y0=lambda x: float(x[0]) - 0.3 * x[1])
        T.set_constraints(dirichlet=lambda t, x: np.sin(t) if x[0] == 0 else None)
        return T, api.couple({'T': T})
    T, sysobj = make(product_api())
    path = sysobj.solve_to_disk(0.05, 0.01, 'run0', parent=str(tmp_path))
    t, y = np.load(f'{path}/t.npy'), np.load(f'{path}/T.npy')
    To, so = make(oracle_api())
    want = []
    for _ in range(t.size):
        so.step(0.01)
        want.append(np.array(To.y, copy=True))
    assert y.shape == (t.size, T.ndim) and np.allclose(t, 0.01 * np.arange(1, t.size + 1))
    assert rel_l2(y, np.array(want)) <= 1e-10




@pytest.mark.parametrize('cname', ['solve', 'user_projection', 'external_field',
                                   'point_reads'])
def test_api_surface_matches_reference_golden(cname):
    """the same API cases against golden vectors produced by the UNMODIFIED reference (tests/golden/api_*.npz; the oracle
    matches them on the CPU, tests/test_oracle_golden.py)"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'api_{cname}.npz'))
    out = cases.API_CASES[cname](product_api())
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.shape(out[k]) == gold[k].shape, k
        assert rel_l2(out[k], gold[k]) <= 1e-10, (cname, k, rel_l2(out[k], gold[k]))


def test_solver_object_facade_steps_like_step():
    """driving obs.integrator by hand (t_bound, status, step()) is the loop fds.step runs (fds.py:284-290)"""
    a, b = heat(product_api()), heat(product_api())
    a.step(0.011)                                         # 5 internal steps of 2e-3 and one of 1e-3
    it = b.integrator
    it.t_bound = 0.011
    n = 0
    while it.status == 'running':
        it.step()
        n += 1
    assert n == 6 and b.t == pytest.approx(a.t, abs=1e-15) and it.t == b.t
    assert np.array_equal(np.asarray(a.y), np.asarray(b.y)) and np.array_equal(it.y[:b.ndim], np.asarray(b.y))
    with pytest.raises(RuntimeError):
        it.step()
    y = it.y
    y[3] += 1.0
    it.y = y                                              # assignment reaches the device state
    assert b.y[3] == pytest.approx(y[3])


def _ensemble_problem(kind):
    """(stepper, fields in engine order, dt) of a fresh small problem"""
    import gds_b200
    api = product_api()
    if kind == 'heat':                      # time-varying Dirichlet values, 3 internal steps per call
        T = heat(api, n=40)
        return T, [T], 0.006
    if kind == 'wave':                      # order 2, collapsed RK4
        a = api.node_gds(cases.g_grid3((9, 10, 11)))
        api.evolve(a, dydt=lambda t, y: a.laplacian(), order=2, max_step=0.1)
        a.set_initial(y0=lambda x: math.exp(-sum((xi - 4.0) ** 2 for xi in x) / 4.0))
        return a, [a], 0.1
    G = cases.g_hex(7, 9)                   # two coupled fields on different domains
    u, c = api.edge_gds(G), api.node_gds(G)
    api.evolve(u, dydt=lambda t, y: 0.1 * u.laplacian(y), max_step=1e-3)
    api.evolve(c, dydt=lambda t, y: -c.advect(u) + 0.05 * c.laplacian(y), max_step=1e-3)
    u.set_initial(y0=lambda e: 0.)
    c.set_initial(y0=lambda x: 0.)
    u.set_constraints(dirichlet=api.zero_edge_bc(cases.boundary_edges_hex_top_bottom(G)))
    sysobj = gds_b200.couple({'u': u, 'c': c})
    return sysobj, [u, c], 2e-3


@pytest.mark.parametrize('kind', ['heat', 'wave', 'hex_coupled'])
def test_ensemble_stepping_equals_the_member_loop_bit_for_bit(kind):
    """step_ensemble (uploads, steps and read-backs of consecutive members overlapped on three streams) against
    write / step / read of one member at a time on a fresh system, two consecutive calls"""
    import torch
    rng = np.random.default_rng(11)
    M = 5
    st, fields, dt = _ensemble_problem(kind)
    sizes = [f.ndim * f._order() for f in fields]
    members = [[rng.standard_normal(n) for n in sizes] for _ in range(M)]

    def pinned(a):
        t = torch.empty(a.size, dtype=torch.float64).pin_memory().numpy()
        t[:] = a
        return t
    ins = [[pinned(a) for a in m] for m in members]
    mid = [[pinned(np.zeros_like(a)) for a in m] for m in members]
    outs = [[pinned(np.zeros_like(a)) for a in m] for m in members]
    single = len(fields) == 1 and not hasattr(st, 'stepper')
    if single:
        st.step_ensemble(dt, [m[0] for m in ins], [m[0] for m in mid])
        st.step_ensemble(dt, [m[0] for m in mid], [m[0] for m in outs])
    else:
        st.step_ensemble(dt, ins, mid)
        st.step_ensemble(dt, mid, outs)
    assert fields[0].t == pytest.approx(2 * dt)
    for k in range(M):
        ref, rf, _ = _ensemble_problem(kind)
        eng = (ref.stepper if hasattr(ref, 'stepper') else ref)._ensure_engine()
        for f, a in zip(rf, members[k]):
            eng.write(f, a)
        ref.step(dt)
        for f, got in zip(rf, mid[k]):
            assert np.array_equal(eng.read(f, full=True), got), (kind, k, 'first call')
        ref.step(dt)
        for f, got in zip(rf, outs[k]):
            assert np.array_equal(eng.read(f, full=True), got), (kind, k, 'second call')
    # the system is left holding the last member's state
    eng = (st.stepper if hasattr(st, 'stepper') else st)._ensure_engine()
    for f, got in zip(fields, outs[-1]):
        assert np.array_equal(eng.read(f, full=True), got)


def test_ensemble_stepping_rejects_what_it_cannot_overlap():
    T = heat(product_api())
    with pytest.raises(ValueError):
        T.step_ensemble(0.002, [np.zeros(T.ndim)], [np.zeros(T.ndim + 1)])
    with pytest.raises(ValueError):
        T.step_ensemble(0.002, [np.zeros(T.ndim, dtype=np.float32)], [np.zeros(T.ndim)])
    # pageable buffers work (the copies serialise)
    a, b = np.linspace(0, 1, T.ndim), np.zeros(T.ndim)
    T.step_ensemble(0.002, [a], [b])
    assert np.array_equal(np.asarray(T.y), b)


def test_fused_cg_iteration_equals_the_six_launch_form_bit_for_bit(monkeypatch):
    """gds_cg_solve runs an iteration in 4 launches (every block totals the partial sums itself, scalars committed through
    shadow slots); GDS_B200_CG_UNFUSED=1 keeps the 6-launch form: same iterates, same iteration count, same bits"""
    import gds_b200
    for make in (lambda: cases.add_weights(cases.g_tri(9, 12), 4), lambda: cases.g_hex(7, 9)):
        u = gds_b200.edge_gds(make())
        y = np.random.default_rng(8).standard_normal(u.ndim)
        monkeypatch.delenv('GDS_B200_CG_UNFUSED', raising=False)
        z1, i1 = u.leray_project(y, return_info=True)
        i1 = dict(i1)
        monkeypatch.setenv('GDS_B200_CG_UNFUSED', '1')
        z2, i2 = u.leray_project(y, return_info=True)
        assert np.array_equal(z1, z2)
        assert all(i1[k] == i2[k] for k in ('iterations', 'relative_residual', 'preconditioner')) and i1['iterations'] > 0
