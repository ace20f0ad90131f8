"""Host-side logic that needs no GPU: fixed-step tables, BC helpers, and the tracer / lowering of evolution laws to
row programs (on a host-only context, which validates programs but never computes)."""
import math

import numpy as np
import pytest

import cases
import gds_b200
from gds_b200 import _lib as L
from gds_b200.engine import step_table
from gds_b200.trace import Lowering, Stage, Time, as_expr, trace_scope, TraceError
from oracle import gds_oracle

OPN = {v: k for k, v in L.OP.items()}


@pytest.fixture(autouse=True)
def host_only(monkeypatch):
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    yield
    gds_b200.Context._default.clear()


@pytest.mark.parametrize('t0,dt,h', [(0., 0.1, 1e-2), (0., 1e-2, 4e-3), (0.3, 1e-2, 1e-2), (0., 0.05, 0.02),
                                      (1.0, 1e-3, 1e-2), (0., 1.0, 0.1)])
def test_step_table_matches_solver_protocol(t0, dt, h):
    """same (t_i, h_i) sequence as the oracle's fixed-step solver object (protocol of gds/ode/midpoint.py:10-32)"""
    used = []

    class Rec(gds_oracle.Euler):
        def _advance(self, t, y, hh):
            used.append((t, hh))
            return y
    sol = Rec(lambda t, y: np.zeros_like(y), t0, np.zeros(1), np.inf, max_step=h)
    t = t0
    for _ in range(5):                       # several step(dt) calls in a row, like fds.step
        sol.t_bound = sol.t + dt
        sol.status = 'running'
        ts, hs, t_new = step_table(t, t + dt, h)
        used.clear()
        while sol.status != 'finished':
            sol.step()
        assert np.array_equal(ts, [g[0] for g in used])
        assert np.array_equal(hs, [g[1] for g in used])
        assert t_new == sol.t
        t = t_new


def test_bc_helpers_follow_the_reference():
    import networkx as nx
    f = gds_b200.combine_bcs({(0, 0): 1.0}, lambda x: 2.0 if x == (1, 1) else None,
                             lambda t, x: t if x in ((0, 0), (2, 2)) else None)
    assert gds_b200.fun_ary(f) == 2
    assert f(5.0, (0, 0)) == 1.0          # static conditions win (boundary.py:21-41)
    assert f(5.0, (1, 1)) == 2.0
    assert f(5.0, (2, 2)) == 5.0
    assert f(5.0, (3, 3)) is None
    g = gds_b200.combine_bcs({(0, 0): 1.0}, lambda x: None)
    assert gds_b200.fun_ary(g) == 1 and g((0, 0)) == 1.0 and g((9, 9)) is None
    dG = nx.Graph([((0, 0), (0, 1))])
    bc = gds_b200.const_edge_bc(dG, 10.)
    assert bc(((0, 0), (0, 1))) == 10. and bc(((0, 1), (0, 0))) == 10. and bc(((5, 5), (5, 6))) is None
    assert gds_b200.zero_edge_bc(dG)(((0, 1), (0, 0))) == 0.
    b = gds_b200.bidict({('a', 'b'): 3})
    assert b[('b', 'a')] == 3 and ('b', 'a') in b and ('a', 'c') not in b


def lower(field, law, scheme='rk4'):
    """trace a law and lower it; returns the list of passes (code as opcode names)"""
    ctx = field._low.ctx
    lw = Lowering(ctx, lambda e: (L.VEC_STAGE if isinstance(e, Stage) else L.VEC_STATE, None, 0), hoist=True)
    with trace_scope():
        root = law(Time(), Stage(field))
    lw.rhs_pass(as_expr(root, field._domain_code, field._low), field._domain_code, field._low, 0, None)
    return [(p.domain, p.when, [OPN[c & 0xff] for c in p.code]) for p in lw.passes]


def test_heat_law_is_one_fused_pass():
    T = gds_b200.node_gds(cases.g_grid(6, 6))
    T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
    T.set_constraints(dirichlet={(0, 0): 1.0})
    passes = lower(T, T.dydt_fun)
    assert passes == [(L.NODES, L.WHEN_STAGE, ['N_LAP', 'MASK_ZERO'])]


def test_accepted_state_subexpressions_are_hoisted_out_of_the_stages():
    T = gds_b200.node_gds(cases.g_grid(6, 6))
    T.set_evolution(dydt=lambda t, y: 0.7 * T.laplacian() + y, max_step=1e-2)     # T.laplacian(): accepted state (App. B.3)
    passes = lower(T, T.dydt_fun)
    assert [p[1] for p in passes] == [L.WHEN_STEP, L.WHEN_STAGE]
    assert 'N_LAP' in passes[0][2] and 'N_LAP' not in passes[1][2]


def test_hodge_laplacian_and_advection_pass_structure():
    G = cases.g_hex(3, 4)
    u = gds_b200.edge_gds(G)
    u.set_evolution(dydt=lambda t, y: 0.1 * u.laplacian(y) - u.advect(y), max_step=1e-3)
    passes = lower(u, u.dydt_fun)
    doms = [p[0] for p in passes]
    # node pass (B1 y), face pass (B2 y), node aggregate pass, then ONE fused edge pass
    assert sorted(doms[:-1]) == sorted([L.NODES, L.FACES, L.NODES]) and doms[-1] == L.EDGES
    last = passes[-1][2]
    assert 'E_GRADT' in last and 'E_CURLT' in last and 'E_ADVFIN' in last


def test_advection_jacobian_vector_product_pass_structure():
    """advect_jvp inside a law (a linearised perturbation equation dv/dt = J(u) v): the aggregates of u and their
    tangent along v are node passes, the rest is fused into edge passes"""
    G = cases.g_hex(3, 4)
    u = gds_b200.edge_gds(G)
    base = np.random.default_rng(0).standard_normal(u.ndim)
    u.set_evolution(dydt=lambda t, y: u.advect_jvp(y, base), max_step=1e-3)
    passes = lower(u, u.dydt_fun)
    # the base state is constant during a step: its aggregates and the S1 factor are hoisted out of the stages;
    # per stage only the tangent aggregates (node pass) and one fused edge pass remain
    assert [(p[0], p[1]) for p in passes] == [(L.NODES, L.WHEN_STEP), (L.EDGES, L.WHEN_STEP),
                                              (L.NODES, L.WHEN_STAGE), (L.EDGES, L.WHEN_STAGE)]
    assert passes[0][2][0] == 'N_EAGG' and 'E_ADVS1' in passes[1][2]
    assert passes[2][2][0] == 'N_EAGGT' and 'E_ADVFIN' in passes[3][2] and 'E_ADVS1' not in passes[3][2]


def test_data_dependent_control_flow_is_rejected():
    T = gds_b200.node_gds(cases.g_grid(4, 4))
    T.set_evolution(dydt=lambda t, y: y if y else -y, max_step=1e-2)
    with pytest.raises(TraceError):
        lower(T, T.dydt_fun)


def test_domain_mismatch_is_rejected():
    G = cases.g_hex(2, 2)
    u, c = gds_b200.edge_gds(G), gds_b200.node_gds(G)
    u.set_evolution(dydt=lambda t, y: y + c.y, max_step=1e-2)
    c.set_evolution(nil=True)
    with pytest.raises(ValueError):
        lower(u, u.dydt_fun)


def test_row_program_validation_in_the_library():
    """gds_pass_run validates programs before touching the device: bad programs fail with a message"""
    import ctypes as C
    T = gds_b200.node_gds(cases.g_grid(4, 4))
    ctx, dev = T._low.ctx, T._low.dev
    d = L.PassDesc()
    code = (C.c_uint32 * 2)(L.OP['ADD'], 0)        # pops an empty stack
    d.domain, d.when, d.n_code, d.code = L.NODES, L.WHEN_STAGE, 1, code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'empty stack' in L.lib.gds_last_error()
    code = (C.c_uint32 * 2)(L.OP['E_GRADT'], 0)    # edge gather in a node pass
    d.code = code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'non-edge pass' in L.lib.gds_last_error()
    # tangent aggregates: the direction vector (operand c) and the output slot are checked like every other operand
    code = (C.c_uint32 * 2)(L.OP['N_EAGGT'] | (0 << 8) | (0 << 16) | (3 << 24), 0)
    d.code = code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'out of range' in L.lib.gds_last_error()
    d.domain = L.EDGES
    code = (C.c_uint32 * 2)(L.OP['E_ADVS1'] | (0 << 8) | (1 << 16), 0)    # aggregates operand missing
    d.code = code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'out of range' in L.lib.gds_last_error()
    d.domain = L.NODES
    code = (C.c_uint32 * 2)(L.OP['E_ADVS1'], 0)                           # edge operator in a node pass
    d.code = code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'non-edge pass' in L.lib.gds_last_error()
    code = (C.c_uint32 * 2)(250, 0)                                        # not an opcode
    d.code = code
    rc = L.lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0)
    assert rc != 0 and b'unknown opcode' in L.lib.gds_last_error()


def kernel_kinds(field, law, scheme='rk4'):
    """kernel family (0 interpreter, 1 Laplacian kernels, 2 term-list kernel) of every pass a law lowers to"""
    import ctypes as C
    from gds_b200.trace import pass_desc
    ctx = field._low.ctx
    lw = Lowering(ctx, lambda e: (L.VEC_STAGE if isinstance(e, Stage) else L.VEC_STATE, None, 0), hoist=True)
    with trace_scope():
        root = law(Time(), Stage(field))
    lw.rhs_pass(as_expr(root, field._domain_code, field._low), field._domain_code, field._low, 0, None)
    kinds = []
    for p in lw.passes:
        d, keep = pass_desc(p)
        k = C.c_int32()
        assert L.lib.gds_pass_kernel_kind(C.byref(d), C.byref(k)) == 0, L.lib.gds_last_error()
        kinds.append(k.value)
    return kinds


def test_kernel_dispatch_of_the_baseline_laws():
    G = cases.g_hex(3, 4)
    T, u = gds_b200.node_gds(G), gds_b200.edge_gds(G)
    T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2)
    T.set_constraints(dirichlet={list(T.X)[0]: 1.0}, neumann={list(T.X)[5]: -0.1})
    u.set_evolution(nil=True)
    assert kernel_kinds(T, lambda t, y: T.laplacian(y)) == [1]                               # C1 heat: Laplacian kernel
    P = gds_b200.node_gds(cases.g_grid(4, 4))
    P.set_evolution(dydt=lambda t, y: P.laplacian(y), max_step=1e-2)
    P.set_constraints(dirichlet={(0, 0): 1.0})
    assert kernel_kinds(P, lambda t, y: 0.7 * P.laplacian(y) - 2.0 * y + 0.5) == [1]         # affine tail stays fused
    assert kernel_kinds(T, lambda t, y: 0.7 * T.laplacian(y) - 2.0 * y) == [1]               # Neumann source + y: two additive vectors
    assert kernel_kinds(T, lambda t, y: 0.7 * T.laplacian(y) - 2.0 * y + T.y) == [0]         # three: interpreter
    # C2 node law: the accepted-state advection term is hoisted to a per-step term-list pass, the stage pass is the
    # Laplacian kernel with two additive vectors (hoisted term + Neumann source)
    assert kernel_kinds(T, lambda t, y: -T.advect(u) + 0.05 * T.laplacian(y)) == [2, 1]
    assert kernel_kinds(T, lambda t, y: -T.advect(u, y) + 0.05 * T.laplacian(y)) == [2]      # stage-form advection: one term-list pass
    assert kernel_kinds(u, lambda t, y: 0.1 * u.laplacian(y)) == [2, 2, 2]                   # C2 edge law: div, curl, edge pass
    assert kernel_kinds(P, lambda t, y: 0.3 * P.laplacian(y) + y - y ** 3) == [1]            # Allen-Cahn: y and y**3 ride along
    assert kernel_kinds(P, lambda t, y: P.laplacian(y ** 3)) == [1]                          # porous medium: chain in the gather
    assert kernel_kinds(P, lambda t, y: P.laplacian(y) + np.sin(y)) == [0]                   # transcendental: interpreter
    assert kernel_kinds(T, lambda t, y: T.laplacian(y) * t) == [0]                           # time-dependent coefficient
    # C4 coupled map lattice: f(y) + eps/8 L f(y) in ONE Laplacian-kernel pass, f fused into the gather as a chain
    f = cases.cml_map()
    x = gds_b200.node_gds(cases.g_grid(5, 5))
    x.set_evolution(map_fun=lambda t, y: f(x, y))
    passes = lower(x, x.map_fun)
    assert len(passes) == 1 and passes[0][2].count('N_LAP') == 1
    assert kernel_kinds(x, x.map_fun) == [1]
    # C5 wave: accepted-state Laplacian hoisted to a per-step Laplacian-kernel pass, the stage pass is pointwise
    a = gds_b200.node_gds(cases.g_grid3((3, 4, 5)))
    a.set_evolution(dydt=lambda t, y: 1.0 * a.laplacian(), order=2, max_step=0.1)
    assert kernel_kinds(a, a.dydt_fun) == [1, 2]


def test_affine_split_of_an_lhs_law():
    """`lhs=` laws are split into a constant part and a part linear in y (fluid.py:26-31 pressure law), including the
    in-place `lhs[rows] = 0.` and `lhs -= ...` of the example"""
    from gds_b200.trace import split_affine, depends_on, MaskZero, Gather, TraceError
    G = cases.g_tri(3, 4)
    pressure, velocity = gds_b200.node_gds(G), gds_b200.edge_gds(G)
    velocity.set_evolution(dydt=lambda t, y: -velocity.advect(), max_step=1e-3)
    rows = np.array([0, 3])

    def pressure_f(t, y):
        lhs = velocity.div(velocity.y / 1e-3 - velocity.advect() + velocity.laplacian() * 0.05)
        lhs[rows] = 0.
        lhs -= pressure.laplacian(y) / 2.0
        return lhs
    pressure.set_evolution(lhs=pressure_f)
    pressure.set_constraints(dirichlet={list(pressure.X)[0]: 1.0})
    with trace_scope():
        root = pressure.lhs_fun(Time(), Stage(pressure))
    const, lin = split_affine(root, pressure)
    assert const is not None and lin is not None
    assert not depends_on(const, pressure) and depends_on(lin, pressure)
    # the constant part carries the row mask of `lhs[rows] = 0.`
    assert isinstance(const, MaskZero) and np.array_equal(const.rows, rows)

    def kinds(e):
        out = [e.kind] if isinstance(e, Gather) else []
        for a in e.args:
            out += kinds(a)
        return out
    assert kinds(lin) == ['N_LAP'] and 'N_BDOT' in kinds(const)
    with trace_scope():
        bad = pressure.laplacian(Stage(pressure)) * Stage(pressure)
    with pytest.raises(TraceError):
        split_affine(bad, pressure)


def _misuses(mod, integrators):
    """the API preconditions of the path (fds.py:72,95,149,189,243,412-414,435; gds.py:168): each entry provokes one"""
    def two_laws():
        T = mod.node_gds(cases.g_grid(3, 3))
        T.set_evolution(dydt=lambda t, y: y, map_fun=lambda t, y: y)

    def no_law():
        mod.node_gds(cases.g_grid(3, 3)).set_evolution()

    def initial_before_law():
        mod.node_gds(cases.g_grid(3, 3)).set_initial(y0=lambda x: 0.)

    def constraints_before_law():
        mod.node_gds(cases.g_grid(3, 3)).set_constraints(dirichlet={(0, 0): 1.})

    def overlap():
        T = mod.node_gds(cases.g_grid(3, 3))
        T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2, integrator=integrators[0])
        T.set_constraints(dirichlet={(0, 0): 1., (0, 1): 2.}, neumann={(0, 1): 0.5})

    def couple_nothing():
        mod.couple({})

    def couple_without_law():
        mod.couple({'T': mod.node_gds(cases.g_grid(3, 3))})

    def couple_late():
        T = mod.node_gds(cases.g_grid(3, 3))
        T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2, integrator=integrators[0])
        T.set_initial(t0=1.0, y0=lambda x: 1.)
        mod.couple({'T': T})

    def incompatible_domains():
        T = mod.node_gds(cases.g_grid(3, 3))
        v = mod.edge_gds(cases.g_grid(4, 4))
        T.advect(v, np.zeros(T.ndim))
    return {'two_laws': two_laws, 'no_law': no_law, 'initial_before_law': initial_before_law,
            'constraints_before_law': constraints_before_law, 'overlap': overlap, 'couple_nothing': couple_nothing,
            'couple_without_law': couple_without_law, 'couple_late': couple_late,
            'incompatible_domains': incompatible_domains}


def test_api_preconditions_raise_like_the_reference():
    """same misuse, same exception type and (where the reference gives one) same message as the oracle, which restates
    the reference's asserts; tests/test_golden_fresh_cpu.py-style check against the reference itself below"""
    ours = _misuses(gds_b200, (gds_b200.Integrators.rk4,))
    theirs = _misuses(gds_oracle, (gds_oracle.Integrators.rk4,))
    for name in ours:
        with pytest.raises(AssertionError) as a:
            ours[name]()
        with pytest.raises(AssertionError) as b:
            theirs[name]()
        ma, mb = str(a.value).split('\n')[0], str(b.value).split('\n')[0]
        if name == 'overlap':       # the message lists the overlapping points: compare up to the list
            ma, mb = ma.split(' on ')[0], mb.split(' on ')[0]
        assert ma == mb, (name, ma, mb)
    T = gds_b200.node_gds(cases.g_grid(3, 3))
    with pytest.raises(ValueError, match='Unsupported integrator type'):         # fds.py:95
        T.set_evolution(dydt=lambda t, y: y, integrator='no such integrator')


@pytest.mark.skipif(__import__('oracle.ref_harness').ref_harness.reference_root() is None, reason='reference not present')
def test_api_preconditions_raise_like_the_unmodified_reference():
    from oracle.ref_harness import load_reference
    ref = load_reference()
    theirs = _misuses(ref, (ref.Integrators.implicit_midpoint,))
    ours = _misuses(gds_b200, (gds_b200.Integrators.rk4,))
    for name in ours:
        with pytest.raises(AssertionError) as a:
            ours[name]()
        with pytest.raises(AssertionError) as b:
            theirs[name]()
        ma, mb = str(a.value).split('\n')[0], str(b.value).split('\n')[0]
        if name == 'overlap':
            ma, mb = ma.split(' on ')[0], mb.split(' on ')[0]
        assert ma == mb, (name, ma, mb)


def test_solver_object_facade_host_side():
    """obs.integrator mirrors the solver-object attributes the reference touches (gds/ode/midpoint.py:10-32): state
    vector with every order block, assignable; t, t_bound, status, max_step; fields without a `dydt=` law have none"""
    T = gds_b200.node_gds(cases.g_grid(4, 5))
    T.set_evolution(dydt=lambda t, y: T.laplacian(), order=2, max_step=2e-2)
    T.set_initial(y0=lambda x: float(x[0]) + 0.1 * x[1])
    it = T.integrator
    assert it is T.integrator and it.status == 'running' and it.t_bound == np.inf and it.max_step == 2e-2
    assert it.t == 0. and it.y.shape == (2 * T.ndim,) and np.array_equal(it.y[:T.ndim], T.y)
    new = np.arange(2 * T.ndim, dtype=float)
    it.y = new
    assert np.array_equal(it.y, new) and np.array_equal(np.asarray(T.y), new[:T.ndim])
    it.y[0] = -1.0                                        # a copy: writes go through the setter only
    assert T.y[0] == 0.0
    M = gds_b200.node_gds(cases.g_grid(3, 3))
    M.set_evolution(map_fun=lambda t, y: 0.5 * y)
    with pytest.raises(AttributeError):
        M.integrator
    # a coupled system has ONE solver object over the concatenated state, in coupling order (fds.py:439-447)
    G = cases.g_hex(2, 2)
    u, c = gds_b200.edge_gds(G), gds_b200.node_gds(G)
    u.set_evolution(dydt=lambda t, y: -y, max_step=4e-3)
    c.set_evolution(dydt=lambda t, y: c.laplacian(y), max_step=1e-3)
    u.set_initial(y0=lambda e: 1.0)
    c.set_initial(y0=lambda x: 2.0)
    st = gds_b200.couple({'u': u, 'c': c}).stepper
    ci = st.integrator
    assert ci.max_step == 1e-3 and ci.t == 0. and ci.y.shape == (u.ndim + c.ndim,)
    assert np.all(ci.y[:u.ndim] == 1.0) and np.all(ci.y[u.ndim:] == 2.0)
    with pytest.raises(Exception):
        u.integrator                      # the coupled system owns the solver object
