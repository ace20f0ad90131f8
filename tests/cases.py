"""Shared case definitions for the parity tests.

Every case is written against an *API adaptor* so that the same definition drives
  * the unmodified reference (tests/golden/make_golden.py, build container only),
  * the CPU oracle (oracle/gds_oracle.py),
  * the CUDA product (gds_b200).
The adaptor supplies the field constructors, `couple`, the BC helpers and `evolve()`, which installs a
fixed-step scheme (the reference needs its integrator object swapped, see make_golden.py).
"""
from __future__ import annotations

import math

import networkx as nx
import numpy as np


# ------------------------------------------------------------------------------------------------
# graphs
# ------------------------------------------------------------------------------------------------
GRAPH_HOOK = None   # the multi-rank tests set this to mark every generated graph as distributed


def _hook(G):
    return GRAPH_HOOK(G) if GRAPH_HOOK is not None else G


def g_grid(m, n, faces_off=True):
    G = nx.grid_2d_graph(m, n)
    if faces_off:
        G.faces = []
    return _hook(G)


def g_hex(m, n):
    return _hook(nx.hexagonal_lattice_graph(m, n))


def g_tri(m, n):
    return _hook(nx.triangular_lattice_graph(m, n))


def g_grid3(dims):
    G = nx.grid_graph(dim=list(dims))
    G.faces = []
    return _hook(G)


def g_rgg(n, r, seed):
    G = nx.random_geometric_graph(n, r, seed=seed)
    G.faces = []
    return _hook(G)


def _gi(obs, x):
    """row of point x in the undistributed field: initial arrays are indexed by it, so that a case gives the same
    problem whether or not the graph is distributed (the reference and the oracle only know `X`)"""
    gi = getattr(obs, 'global_index', None)
    return gi(x) if gi is not None else obs.X[x]


def add_weights(G, seed, lo=0.5, hi=2.0):
    rng = np.random.default_rng(seed)
    for (u, v) in G.edges():
        G[u][v]['weight'] = float(rng.uniform(lo, hi))
    return G


def boundary_edges_hex_top_bottom(G):
    """Sub-graph holding the top and bottom boundary edges of a hexagonal lattice (for zero_edge_bc)."""
    ys = [p[1] for p in nx.get_node_attributes(G, 'pos').values()]
    lo, hi = min(ys), max(ys)
    h = math.sqrt(3) / 2
    pos = nx.get_node_attributes(G, 'pos')
    keep = [n for n in G.nodes() if pos[n][1] < lo + 0.6 * h or pos[n][1] > hi - 0.6 * h]
    return G.subgraph(keep).copy()


def tri_boundaries(G):
    """(lid, walls): top-row edges, and left/right/bottom boundary edges of a triangular lattice."""
    pos = nx.get_node_attributes(G, 'pos')
    xs = [p[0] for p in pos.values()]
    ys = [p[1] for p in pos.values()]
    x0, x1, y0, y1 = min(xs), max(xs), min(ys), max(ys)
    top = [n for n in G.nodes() if pos[n][1] > y1 - 1e-9]
    bottom = [n for n in G.nodes() if pos[n][1] < y0 + 1e-9]
    left = [n for n in G.nodes() if pos[n][0] < x0 + 0.51]
    right = [n for n in G.nodes() if pos[n][0] > x1 - 0.51]
    lid = G.subgraph(top).copy()
    walls = nx.compose_all([G.subgraph(bottom).copy(), G.subgraph(left).copy(), G.subgraph(right).copy()])
    return lid, walls


# ------------------------------------------------------------------------------------------------
# operator cases: graph + seeded inputs -> dict of operator outputs
# ------------------------------------------------------------------------------------------------
OP_GRAPHS = {
    'grid_6x7': lambda: g_grid(6, 7),
    'hex_3x4': lambda: g_hex(3, 4),
    'hex_4x5_w': lambda: add_weights(g_hex(4, 5), 11),
    'tri_4x6': lambda: g_tri(4, 6),
    'tri_5x7_w': lambda: add_weights(g_tri(5, 7), 12),
    'rgg_60': lambda: g_rgg(60, 0.25, 3),
}


def operator_outputs(api, gname):
    """Apply every hot-path operator to seeded inputs; returns {key: ndarray} plus the index maps."""
    G = OP_GRAPHS[gname]()
    rng = np.random.default_rng(1234)
    out = {}
    nd = api.node_gds(G)
    ed = api.edge_gds(G)
    V, E = nd.ndim, ed.ndim
    yv, ye = rng.standard_normal(V), rng.standard_normal(E)
    ye[rng.integers(0, E, size=max(1, E // 10))] = 0.0  # exact zeros exercise sign()==0 branches
    out['in_yv'], out['in_ye'] = yv, ye
    nodes_list = list(nd.X.keys())
    edges_list = list(ed.X.keys())
    out['map_edges'] = np.array([[nd.X[a], nd.X[b]] for (a, b) in edges_list], dtype=np.int64).reshape(E, 2)
    out['map_nodes_repr'] = np.array([repr(n) for n in nodes_list])
    out['edge_weights'] = np.asarray(ed.edge_weights, dtype=float)
    # node field: Dirichlet on a few nodes, Neumann on a few others
    d_nodes = {nodes_list[i]: float(0.25 * i) for i in range(0, V, 7)}
    n_nodes = {nodes_list[i]: float(-0.1 * (i % 5 + 1)) for i in range(3, V, 11) if nodes_list[i] not in d_nodes}
    api.evolve(nd, dydt=lambda t, y: nd.laplacian(y), max_step=1e-2)
    nd.set_constraints(dirichlet=d_nodes, neumann=n_nodes)
    out['node_grad'] = nd.grad(yv)
    out['node_laplacian'] = nd.laplacian(yv)
    out['node_bilaplacian'] = nd.bilaplacian(yv)
    out['node_advect'] = nd.advect(ye, yv)
    out['node_lie_advect'] = nd.lie_advect(ye, yv)
    out['edge_leray'] = ed.leray_project(ye)      # before any Dirichlet condition (gds.py:418-419)
    # edge field: Dirichlet on a few edges
    d_edges = {edges_list[i]: float(0.1 * i) for i in range(0, E, 9)}
    api.evolve(ed, dydt=lambda t, y: ed.laplacian(y), max_step=1e-2)
    ed.set_constraints(dirichlet=d_edges)
    free = np.array(sorted(set(range(2, E, 13))), dtype=np.intp)
    normal = np.array(sorted(set(range(5, E, 17))), dtype=np.intp)
    out['edge_div'] = ed.div(ye)
    out['edge_influx'] = np.asarray(ed.influx(ye)).ravel()
    out['edge_outflux'] = np.asarray(ed.outflux(ye)).ravel()
    out['edge_dd'] = ed.dd_(ye)
    out['edge_dd_normal'] = ed.dd_(ye, normal=normal)
    out['edge_advect'] = ed.advect(ye)
    nF = len(ed.faces)
    out['n_faces'] = np.array([nF])
    if nF > 0:
        faces_list = list(ed.faces.keys())
        out['map_faces_ptr'] = np.cumsum([0] + [len(f) for f in faces_list]).astype(np.int64)
        out['map_faces_nodes'] = np.array([nd.X[n] for f in faces_list for n in f], dtype=np.int64)
        out['edge_curl'] = ed.curl(ye)
        out['edge_d_d'] = ed.d_d(ye)
        out['edge_laplacian'] = ed.laplacian(ye)
        out['edge_laplacian_fn'] = ed.laplacian(ye, free=free, normal=normal)
        out['edge_bilaplacian'] = ed.bilaplacian(ye)
        fd = api.face_gds(G)
        yf = rng.standard_normal(fd.ndim)
        out['in_yf'] = yf
        out['face_laplacian'] = fd.laplacian(yf)
    return {k: np.asarray(v) for k, v in out.items()}


JAC_GRAPHS = ('hex_3x4', 'tri_4x6', 'hex_4x5_w')


def jacobian_outputs(api, gname):
    """Jacobian of the edge self-advection at a seeded state (gds.py:391-395: autograd through advect_torch; float64
    state, the reference's float32 operator copies)."""
    import torch
    G = OP_GRAPHS[gname]()
    ed = api.edge_gds(G)
    y = np.random.default_rng(5).standard_normal(ed.ndim)
    y[::7] = 0.0                                   # zero-flow edges: sign() == 0 rows and columns
    J = ed.advect_jac(torch.from_numpy(y.copy()))
    return {'in_y': y, 'jac': np.asarray(J.numpy() if hasattr(J, 'numpy') else J, dtype=float)}


# ------------------------------------------------------------------------------------------------
# trajectory cases
# ------------------------------------------------------------------------------------------------
def _snap(sysobj, observables, dt, nsteps, every):
    """step `nsteps` times by `dt`; record each observable's y every `every` steps (and at the end)."""
    rec = {k: [] for k in observables}
    rec_t = []
    for i in range(1, nsteps + 1):
        sysobj.step(dt)
        if i % every == 0 or i == nsteps:
            for k, o in observables.items():
                rec[k].append(np.array(o.y, dtype=float, copy=True))
            rec_t.append(float(sysobj.t))
    out = {f'y_{k}': np.array(v) for k, v in rec.items()}
    out['t'] = np.array(rec_t)
    for k, o in observables.items():
        part = getattr(o, 'partition', None)
        if part is not None:                      # one rank of a distributed field
            out[f'gid_{k}'] = np.asarray(o.global_ids)
            out[f'nown_{k}'] = np.array([part.n_own])
            out[f'own_{k}'] = np.asarray(o.owned_mask)
    return out


def case_heat_c1(api, n=100, nsteps=1000, every=250, scheme='rk4'):
    """BASELINE config 1: heat equation, grid n x n, RK4 h=1e-2, static + time-varying Dirichlet
    (README.md:208-211 / examples/heat.py:56-63)."""
    G = g_grid(n, n)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=1e-2, scheme=scheme)
    c = n / 2
    T.set_initial(y0=lambda x: math.exp(-((x[0] - c) ** 2 + (x[1] - c) ** 2) / (n / 2)))
    T.set_constraints(dirichlet=api.combine_bcs(
        lambda x: 0. if x[0] == n - 1 else None,
        lambda t, x: math.sin(t + x[1] / 4) ** 2 if x[0] == 0 else None))
    return _snap(T, {'T': T}, 1e-2, nsteps, every)


def case_heat_mixed(api, scheme='rk4'):
    """Dirichlet + Neumann heat equation on a weighted 12x9 grid (examples/heat.py:25-39 pattern);
    accepted-state form `T.laplacian()` (no argument: Appendix B.3) and step(dt) spanning several max_steps."""
    G = add_weights(g_grid(12, 9), 5)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: 0.7 * T.laplacian(), max_step=4e-3, scheme=scheme)
    rng = np.random.default_rng(7)
    y0 = rng.random(12 * 9)
    T.set_initial(y0=lambda x: y0[x[0] * 9 + x[1]])     # = y0[T.X[x]] on the undistributed graph

    def dirichlet(t, x):
        if x[0] == 0 or x[0] == 11:
            return 0.5
        if x[1] == 0:
            return 1.0 + 0.1 * t
        return None

    def neumann(t, x):
        if x[1] == 8 and x[0] not in (0, 11):
            return -0.1
        return None
    T.set_constraints(dirichlet=dirichlet, neumann=neumann)
    return _snap(T, {'T': T}, 1e-2, 40, 10)   # dt = 2.5 max_steps -> 3 internal steps, the last one shorter


def case_reaction_diffusion(api):
    """stage-argument form with pointwise nonlinearity: dy/dt = 0.3*L y + y - y**3 (Allen-Cahn pattern)."""
    G = g_tri(6, 9)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: 0.3 * T.laplacian(y) + y - y ** 3, max_step=5e-3)
    rng = np.random.default_rng(21)
    y0 = rng.uniform(-1, 1, T.ndim)
    T.set_initial(y0=lambda x: y0[T.X[x]])
    return _snap(T, {'T': T}, 5e-3, 60, 20)


def case_hex_advdiff_c2(api, m=6, n=8, nsteps=50, every=25):
    """BASELINE config 2 at small size: edge diffusion (Hodge-1 Laplacian, Dirichlet 0 on top/bottom edges)
    coupled with node advection-diffusion (Neumann on top row), RK4 h=1e-3."""
    G = g_hex(m, n)
    u = api.edge_gds(G)
    c = api.node_gds(G)
    nu, kappa = 0.1, 0.05
    api.evolve(u, dydt=lambda t, y: nu * u.laplacian(y), max_step=1e-3)
    api.evolve(c, dydt=lambda t, y: -c.advect(u) + kappa * c.laplacian(y), max_step=1e-3)
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    u0 = 0.1 * rng0.standard_normal(G.number_of_edges())
    c0 = rng1.random(G.number_of_nodes())
    u.set_initial(y0=lambda e: u0[_gi(u, e)])
    c.set_initial(y0=lambda x: c0[_gi(c, x)])
    u.set_constraints(dirichlet=api.zero_edge_bc(boundary_edges_hex_top_bottom(G)))
    pos = nx.get_node_attributes(G, 'pos')
    ytop = max(p[1] for p in pos.values())
    c.set_constraints(neumann=lambda x: -0.1 if pos[x][1] > ytop - 1e-9 else None)
    sysobj = api.couple({'u': u, 'c': c})
    return _snap(sysobj, {'u': u, 'c': c}, 1e-3, nsteps, every)


def case_ns_cavity_c3(api, m=6, n=10, nsteps=40, every=20):
    """BASELINE config 3 at small size: Navier-Stokes velocity law (examples/fluid.py:33-34) on a triangular
    lattice with the pressure held as a given node vector; lid const_edge_bc (time-varying form), no-slip walls
    (examples/lid_driven_cavity.py:17-27).  Accepted-state operators (no y argument), RK4 h=1e-4."""
    G = g_tri(m, n)
    p = api.node_gds(G)
    u = api.edge_gds(G)
    rho, visc = 0.1, 200.0
    rngp = np.random.default_rng(3)
    p0 = 0.01 * rngp.standard_normal(G.number_of_nodes())
    api.evolve_nil(p)
    p.set_initial(y0=lambda x: p0[_gi(p, x)])
    api.evolve(u, dydt=lambda t, y: -u.advect() - p.grad() / rho + u.laplacian() * visc / rho, max_step=1e-4)
    rngu = np.random.default_rng(4)
    u0 = 0.05 * rngu.standard_normal(G.number_of_edges())
    u.set_initial(y0=lambda e: u0[_gi(u, e)])
    lid, walls = tri_boundaries(G)
    lid_bc = api.const_edge_bc(lid, 10.0)
    u.set_constraints(dirichlet=api.combine_bcs(api.zero_edge_bc(walls), lambda t, e: lid_bc(e)))
    sysobj = api.couple({'u': u, 'p': p})
    return _snap(sysobj, {'u': u}, 1e-4, nsteps, every)


def case_ns_pressure(api, m=5, n=8, nsteps=12, every=6, rho=1.0, visc=0.05, h=1e-3):
    """Navier-Stokes with the pressure solved every step (examples/fluid.py:14-38 `navier_stokes`): the pressure field
    is defined by `lhs=` -- divergence of the provisional velocity minus the pressure Laplacian, set to zero -- with a
    pressure drop between an inlet and an outlet node; the velocity law reads `pressure.grad()`.  NOT in the golden
    set of round 1; pinned since (tests/golden/lhs_ns_pressure.npz: the reference's cvx path with oracle/cvx_shim.py as its CVXPY).  h equals the law's
    min_step so that `max(min_step, velocity.dt)` does not depend on the reference's `dt` bookkeeping (App. B.7)."""
    G = g_tri(m, n)
    pressure, velocity = api.node_gds(G), api.edge_gds(G)
    nodes = list(pressure.X.keys())
    inlet, outlet = nodes[0], nodes[-1]
    v_free = np.array([pressure.X[inlet], pressure.X[outlet]], dtype=np.intp)
    min_step = 1e-3

    def pressure_f(t, y):
        dt = max(min_step, velocity.dt)
        lhs = velocity.div(velocity.y / dt - velocity.advect() + velocity.laplacian() * visc / rho)
        lhs[v_free] = 0.
        lhs -= pressure.laplacian(y) / rho
        return lhs

    def velocity_f(t, y):
        return -velocity.advect() - pressure.grad() / rho + velocity.laplacian() * visc / rho
    api.evolve_lhs(pressure, pressure_f)
    api.evolve(velocity, dydt=velocity_f, max_step=h)
    u0 = 0.05 * np.random.default_rng(13).standard_normal(velocity.ndim)
    velocity.set_initial(y0=lambda e: u0[velocity.X[e]])
    pressure.set_constraints(dirichlet={inlet: 1.0, outlet: -1.0})
    sysobj = api.couple({'velocity': velocity, 'pressure': pressure})
    return _snap(sysobj, {'velocity': velocity, 'pressure': pressure}, h, nsteps, every)


def case_wave_c5(api, dims=(6, 7, 8), nsteps=50, every=25):
    """BASELINE config 5 at small size: wave equation on a 3-D grid, order=2, accepted-state Laplacian
    (examples/wave.py:14), RK4 h=0.1, no Dirichlet."""
    G = g_grid3(dims)
    a = api.node_gds(G)
    c2 = 1.0
    api.evolve(a, dydt=lambda t, y: c2 * a.laplacian(), order=2, max_step=0.1)
    ctr = [(d - 1) / 2 for d in dims[::-1]]
    a.set_initial(y0=lambda x: math.exp(-sum((xi - ci) ** 2 for xi, ci in zip(x, ctr)) / 4.0))
    return _snap(a, {'a': a}, 0.1, nsteps, every)


def cml_map(eps=0.3, deg=8.0):
    return lambda x, y: (1 - 1.7 * y ** 2) + (eps / deg) * x.laplacian(1 - 1.7 * y ** 2)


def case_cml_c4(api, n=400, iters=30, every=10):
    """BASELINE config 4 at small size: coupled map lattice on a random geometric graph, map mode."""
    r = math.sqrt(8 / (math.pi * n))
    G = g_rgg(n, r, 42)
    x = api.node_gds(G)
    f = cml_map()
    api.evolve_map(x, map_fun=lambda t, y: f(x, y), dt=1.0)
    y0 = np.random.default_rng(2).uniform(-1, 1, n)
    x.set_initial(y0=lambda v: y0[v])                    # nodes are 0..n-1: = y0[x.X[v]] on the undistributed graph
    return _snap(x, {'x': x}, 1.0, iters, every)


TRAJ_CASES = {
    'heat_c1': case_heat_c1,
    'heat_c1_euler_small': lambda api: case_heat_c1(api, n=20, nsteps=100, every=50, scheme='euler'),
    'heat_mixed': case_heat_mixed,
    'reaction_diffusion': case_reaction_diffusion,
    'hex_advdiff_c2': case_hex_advdiff_c2,
    'ns_cavity_c3': case_ns_cavity_c3,
    'wave_c5': case_wave_c5,
    'cml_c4': case_cml_c4,
}


# ------------------------------------------------------------------------------------------------
# `lhs=` fields (gds/fds.py:97-113,196-202,307-328).  Golden vectors in tests/golden/lhs_*.npz come from the UNMODIFIED
# reference with oracle/cvx_shim.py standing in for CVXPY (absent here): the reference assembles the programme, the
# stand-in solves it exactly.
# ------------------------------------------------------------------------------------------------
def case_poisson_lhs(api, m=6, n=7, nsteps=3):
    """a single `lhs=` field stepped on its own (fds.step -> step_cvx): weighted Poisson problem L0 y = f with
    time-varying Dirichlet values on two rows of the grid"""
    G = add_weights(g_grid(m, n), 21)
    p = api.node_gds(G)
    f = np.random.default_rng(22).standard_normal(p.ndim)
    api.evolve_lhs(p, lambda t, y: p.laplacian(y) - f)
    p.set_constraints(dirichlet=lambda t, x: (1.0 + t) * x[1] if x[0] == 0 else (-0.5 if x[0] == m - 1 else None))
    ys, ts = [], []
    for _ in range(nsteps):
        p.step(0.25)
        ys.append(np.array(p.y, dtype=float, copy=True))
        ts.append(float(p.t))
    return {'y_p': np.array(ys), 't': np.array(ts)}


def case_stokes_hex_lhs(api, m=4, n=5, nsteps=8, every=4, rho=0.7, visc=0.3, h=1e-3):
    """Stokes flow (examples/fluid.py:40-41: navier_stokes with the advection switched off) on a WEIGHTED hexagonal
    lattice, inlet and outlet made of several nodes with a time-varying pressure drop, no-slip edges top and bottom"""
    G = add_weights(g_hex(m, n), 23)
    pressure, velocity = api.node_gds(G), api.edge_gds(G)
    pos = nx.get_node_attributes(G, 'pos')
    xs = [q[0] for q in pos.values()]
    x0, x1 = min(xs), max(xs)
    inlet = [x for x in pressure.X if pos[x][0] < x0 + 1e-9]
    outlet = [x for x in pressure.X if pos[x][0] > x1 - 1e-9]
    v_free = np.array([pressure.X[x] for x in inlet + outlet], dtype=np.intp)
    min_step = 1e-3

    def pressure_f(t, y):
        dt = max(min_step, velocity.dt)
        lhs = velocity.div(velocity.y / dt + velocity.laplacian() * visc / rho)
        lhs[v_free] = 0.
        lhs -= pressure.laplacian(y) / rho
        return lhs
    api.evolve_lhs(pressure, pressure_f)
    api.evolve(velocity, dydt=lambda t, y: -pressure.grad() / rho + velocity.laplacian() * visc / rho, max_step=h)
    u0 = 0.05 * np.random.default_rng(24).standard_normal(velocity.ndim)
    velocity.set_initial(y0=lambda e: u0[velocity.X[e]])
    velocity.set_constraints(dirichlet=api.zero_edge_bc(boundary_edges_hex_top_bottom(G)))
    ins, outs = set(inlet), set(outlet)
    pressure.set_constraints(dirichlet=lambda t, x: (1.0 + 40.0 * t) if x in ins else (-1.0 if x in outs else None))
    sysobj = api.couple({'velocity': velocity, 'pressure': pressure})
    return _snap(sysobj, {'velocity': velocity, 'pressure': pressure}, h, nsteps, every)


LHS_CASES = {
    'ns_pressure': lambda api: case_ns_pressure(api),
    'poisson': case_poisson_lhs,
    'stokes_hex_w': case_stokes_hex_lhs,
}


# ------------------------------------------------------------------------------------------------
# corners of the reference semantics and operator combinations the BASELINE
# cases do not reach; small, golden vectors from the unmodified reference in tests/golden/quirk_*.npz
# ------------------------------------------------------------------------------------------------
def quirk_order2_dirichlet(api):
    """Appendix B.4: for order=2 the Dirichlet overwrite lands on the derivative block (fds.py:196,362)"""
    G = g_grid(7, 6)
    a = api.node_gds(G)
    api.evolve(a, dydt=lambda t, y: 0.8 * a.laplacian(), order=2, max_step=0.05)
    a.set_initial(y0=lambda x: math.sin(x[0]) * math.cos(x[1]))
    a.set_constraints(dirichlet=lambda t, x: 0.3 * math.cos(t) if x[0] == 0 else None)
    out = {}
    for i in range(3):
        a.step(0.1)
        out[f'y{i}'] = np.array(a.y, copy=True)
    return out


def quirk_map_time_varying_dirichlet(api):
    G = g_grid(6, 5)
    x = api.node_gds(G)
    api.evolve_map(x, map_fun=lambda t, y: 0.5 * y + 0.1 * x.laplacian(y) + 0.01 * t, dt=1.0)
    x.set_initial(y0=lambda p: 0.1 * p[0] - 0.05 * p[1])
    x.set_constraints(dirichlet=lambda t, p: math.sin(0.3 * t) if p[1] == 0 else None)
    out = {}
    for i in range(5):
        x.step(1.0)
        out[f'y{i}'] = np.array(x.y, copy=True)
    return out


def quirk_euler_coupled_max_steps(api):
    """coupled_fds uses the smallest max_step of its systems (fds.py:437)"""
    G = g_hex(3, 3)
    u, c = api.edge_gds(G), api.node_gds(G)
    api.evolve(u, dydt=lambda t, y: -0.2 * y + 0.05 * u.laplacian(y), max_step=4e-3, scheme='euler')
    api.evolve(c, dydt=lambda t, y: -c.advect(u, y) + 0.1 * u.div(), max_step=1e-3, scheme='euler')
    r = np.random.default_rng(5)
    u0, c0 = r.standard_normal(u.ndim), r.random(c.ndim)
    u.set_initial(y0=lambda e: u0[u.X[e]])
    c.set_initial(y0=lambda x: c0[c.X[x]])
    s = api.couple({'u': u, 'c': c})
    out = {}
    for i in range(3):
        s.step(2.5e-3)
        out[f'u{i}'], out[f'c{i}'] = np.array(u.y, copy=True), np.array(c.y, copy=True)
    out['t'] = np.array([s.t])
    return out


def quirk_edge_free_normal_fluxes(api):
    G = add_weights(g_tri(4, 6), 8)
    u, p = api.edge_gds(G), api.node_gds(G)
    E = u.ndim
    free, normal = np.arange(1, E, 7), np.arange(3, E, 5)
    api.evolve_nil(p)
    p.set_initial(y0=lambda x: 0.2 * x[0] - 0.1 * x[1])
    api.evolve(u, dydt=lambda t, y: 0.3 * u.laplacian(y, free=free, normal=normal) - 0.5 * p.grad() + 0.05 * u.dd_(y),
               max_step=2e-3)
    u0 = np.random.default_rng(6).standard_normal(E)
    u.set_initial(y0=lambda e: u0[u.X[e]])
    u.set_constraints(dirichlet={list(u.X.keys())[2]: 0.7})
    s = api.couple({'u': u, 'p': p})
    for _ in range(4):
        s.step(2e-3)
    return {'u': np.array(u.y, copy=True), 'influx': np.asarray(u.influx()).ravel(),
            'outflux': np.asarray(u.outflux()).ravel(), 'curl': np.asarray(u.curl())}


def quirk_lie_advect_bilaplacian(api):
    G = g_grid(6, 7)
    v, T = api.edge_gds(G), api.node_gds(G)
    api.evolve_nil(v)
    v0 = np.random.default_rng(7).standard_normal(v.ndim)
    v.set_initial(y0=lambda e: v0[v.X[e]])
    api.evolve(T, dydt=lambda t, y: -0.4 * T.lie_advect(v, y) - 0.01 * T.bilaplacian(y) + 0.2 * T.laplacian(y),
               max_step=1e-3)
    T.set_initial(y0=lambda x: math.exp(-((x[0] - 3) ** 2 + (x[1] - 3) ** 2) / 4))
    T.set_constraints(dirichlet={(0, 0): 0.0}, neumann={(5, 6): 0.3})
    for _ in range(4):
        T.step(1e-3)
    return {'T': np.array(T.y, copy=True)}


def quirk_face_diffusion(api):
    G = g_hex(3, 4)
    w = api.face_gds(G)
    api.evolve(w, dydt=lambda t, y: 0.2 * w.laplacian(y) - 0.1 * y, max_step=5e-3)
    w0 = np.random.default_rng(8).standard_normal(w.ndim)
    w.set_initial(y0=lambda f: w0[w.X[f]])
    for _ in range(3):
        w.step(5e-3)
    return {'w': np.array(w.y, copy=True)}


def quirk_accepted_vs_stage(api):
    """Appendix B.3: `T.laplacian()` (accepted state) and `T.laplacian(y)` (stage state) are different laws"""
    out = {}
    for stage in (False, True):
        G = g_grid(4, 4)
        T = api.node_gds(G)
        api.evolve(T, dydt=(lambda t, y: T.laplacian(y)) if stage else (lambda t, y: T.laplacian()), max_step=5e-2)
        T.set_initial(y0=lambda x: float(x == (1, 1)))
        for _ in range(10):
            T.step(5e-2)
        out['stage' if stage else 'accepted'] = np.array(T.y, copy=True)
    return out


def quirk_map_coupled_continuous(api):
    """fds.py:491-507: a recurrence map coupled with a continuous field is stepped, before EVERY right-hand-side evaluation
    of the shared integrator, to that evaluation's stage time; it is applied when its own clock has run `dt` since the last
    application -- in the middle of an RK4 step if that is where the stage time falls -- reads the ACCEPTED continuous
    state, and the continuous law sees the new map state from that stage on."""
    G = g_grid(6, 5)
    T, m = api.node_gds(G), api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: T.laplacian(y) + 0.5 * m.y, max_step=1e-2)
    api.evolve_map(m, map_fun=lambda t, y: 0.9 * y + 0.2 * T.y + 0.01 * t, dt=0.025)
    T.set_initial(y0=lambda x: math.sin(x[0]) + 0.1 * x[1])
    m.set_initial(y0=lambda x: 0.05 * x[0] - 0.02 * x[1])
    T.set_constraints(dirichlet=lambda t, x: math.cos(t) if x[0] == 0 else None)
    sysobj = api.couple({'T': T, 'm': m})
    out = {}
    for i in range(6):
        sysobj.step(0.02)
        out[f'T{i}'], out[f'm{i}'] = np.array(T.y, copy=True), np.array(m.y, copy=True)
    out['t'] = np.array([sysobj.t])
    return out


QUIRK_CASES = {
    'map_coupled_continuous': quirk_map_coupled_continuous,
    'order2_dirichlet': quirk_order2_dirichlet,
    'map_time_varying_dirichlet': quirk_map_time_varying_dirichlet,
    'euler_coupled_max_steps': quirk_euler_coupled_max_steps,
    'edge_free_normal_fluxes': quirk_edge_free_normal_fluxes,
    'lie_advect_bilaplacian': quirk_lie_advect_bilaplacian,
    'face_diffusion': quirk_face_diffusion,
    'accepted_vs_stage': quirk_accepted_vs_stage,
}


# ------------------------------------------------------------------------------------------------
# the rest of the drop-in surface (System.solve, project=, fields outside the system, point reads, derived observables,
# single-edge differences); golden vectors from the unmodified reference in tests/golden/api_*.npz
# ------------------------------------------------------------------------------------------------
def _api_heat(api, n=10):
    G = g_grid(n, n)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=2e-3)
    T.set_initial(y0=lambda x: math.exp(-((x[0] - 4) ** 2 + (x[1] - 5) ** 2) / 6))
    T.set_constraints(dirichlet=lambda t, x: math.cos(t) if x[0] == 0 else None)
    return T


def api_solve(api):
    """System.solve (system.py:38-57): states after every call of step(dt), initial state first"""
    T = _api_heat(api)
    t, data = api.couple({'T': T}).solve(0.02, 0.004)
    return {'t': np.asarray(t, dtype=float), 'T': np.asarray(data['T'], dtype=float)}


def api_user_projection(api):
    """project= is arbitrary host code applied after every step (fds.py:363)"""
    T = _api_heat(api, n=8)
    T.set_constraints(dirichlet={(0, 0): 1.0}, project=lambda y: np.clip(y, 0.0, 0.6))
    T.step(0.01)
    return {'T': np.array(T.y, copy=True)}


def api_external_field(api):
    """a law may read a field outside its system; step(dt) spans several max_steps"""
    G = g_grid(7, 9)
    src, T = api.node_gds(G), api.node_gds(G)
    api.evolve_nil(src)
    src.set_initial(y0=lambda x: 0.1 * x[0] - 0.05 * x[1])
    api.evolve(T, dydt=lambda t, y: T.laplacian(y) + 2.0 * src.y - y * t, max_step=3e-3)
    T.set_initial(y0=lambda x: float(x[0] == 3))
    for _ in range(3):
        T.step(1e-2)
    return {'T': np.array(T.y, copy=True), 't': np.array([T.t])}


def api_point_reads(api):
    """oriented point reads (gds.py:273-274), curl of the evolved field, single-edge difference (gds.py:149-150)"""
    G = add_weights(g_hex(3, 4), 2)
    u, c = api.edge_gds(G), api.node_gds(G)
    rng = np.random.default_rng(3)
    y0 = rng.standard_normal(u.ndim)
    api.evolve(u, dydt=lambda t, y: 0.1 * u.laplacian(y), max_step=1e-3)
    u.set_initial(y0=lambda e: y0[u.X[e]])
    u.step(3e-3)
    e = list(u.X.keys())[5]
    x = rng.standard_normal(c.ndim)
    api.evolve_nil(c)
    c.set_initial(y0=lambda n: x[c.X[n]])
    return {'read': np.array([u(e), u((e[1], e[0]))]), 'curl': np.asarray(u.curl(), dtype=float),
            'partial': np.array([c.partial(e)]), 'y': np.array(u.y, copy=True)}


API_CASES = {'solve': api_solve, 'user_projection': api_user_projection, 'external_field': api_external_field,
             'point_reads': api_point_reads}


# ------------------------------------------------------------------------------------------------
# empty / tiny / ragged graphs, sizes off the 32- and 128-row boundaries; golden vectors from the unmodified reference
# in tests/golden/ragged_*.npz for every graph the reference can construct
# ------------------------------------------------------------------------------------------------
def _nofaces(G):
    G.faces = []
    return G


def star(k, seed=None):
    G = nx.star_graph(k)
    if seed is not None:
        rng = np.random.default_rng(seed)
        for u, v in G.edges():
            G[u][v]['weight'] = float(rng.uniform(0.5, 2.0))
    return _nofaces(G)


RAGGED = {
    'single_node': lambda: _nofaces(nx.empty_graph(1)),
    'isolated_nodes': lambda: _nofaces(nx.empty_graph(5)),
    'single_edge': lambda: _nofaces(nx.path_graph(2)),
    'path_33': lambda: _nofaces(nx.path_graph(33)),
    'path_129': lambda: _nofaces(nx.path_graph(129)),
    'cycle_127': lambda: _nofaces(nx.cycle_graph(127)),
    'star_200': lambda: star(200),
    'star_70_weighted': lambda: star(70, seed=3),
    'complete_40': lambda: _nofaces(nx.complete_graph(40)),
    'grid_5x26_plus_isolated': lambda: _nofaces(nx.disjoint_union(nx.convert_node_labels_to_integers(nx.grid_2d_graph(5, 26)),
                                                                  nx.empty_graph(3))),
}


def run_ops(api, G):
    rng = np.random.default_rng(99)
    nd, ed = api.node_gds(G), api.edge_gds(G)
    V, E = nd.ndim, ed.ndim
    yv, ye = rng.standard_normal(V), rng.standard_normal(E)
    if E:
        ye[::5] = 0.0
    out = {'lap': nd.laplacian(yv), 'bilap': nd.bilaplacian(yv)}
    if E:
        out.update(grad=nd.grad(yv), nadv=nd.advect(ye, yv), lie=nd.lie_advect(ye, yv), div=ed.div(ye),
                   influx=np.asarray(ed.influx(ye)).ravel(), outflux=np.asarray(ed.outflux(ye)).ravel(),
                   dd=ed.dd_(ye), eadv=ed.advect(ye), elap=ed.laplacian(ye))
    # a few RK4 steps of a nonlinear law with a Dirichlet node
    api.evolve(nd, dydt=lambda t, y: nd.laplacian(y) - 0.1 * y ** 3, max_step=1e-3)
    nd.set_initial(y0=lambda x: yv[nd.X[x]])
    first = next(iter(nd.X))
    nd.set_constraints(dirichlet={first: 0.5})
    for _ in range(3):
        nd.step(2.5e-3)
    out['traj'] = np.array(nd.y, copy=True)
    return out


def fuzz_graphs():
    """random / irregular graphs: components, isolated nodes, leaves, hubs, weights, shuffled labels"""
    rng = np.random.default_rng(17)

    def weighted(G, seed):
        r = np.random.default_rng(seed)
        for u, v in G.edges():
            G[u][v]['weight'] = float(r.uniform(0.3, 3.0))
        return G
    out = {
        'gnp_40': nx.gnp_random_graph(40, 0.12, seed=4),                       # isolated nodes, several components
        'gnp_60_weighted': weighted(nx.gnp_random_graph(60, 0.08, seed=5), 1),
        'regular_3_50': nx.random_regular_graph(3, 50, seed=6),
        'barbell': nx.barbell_graph(7, 4),
        'wheel_weighted': weighted(nx.wheel_graph(23), 2),
        'ladder': nx.ladder_graph(19),
        'tree': nx.bfs_tree(nx.gnp_random_graph(45, 0.2, seed=8), 0).to_undirected(),
        'shuffled_labels': nx.relabel_nodes(nx.cycle_graph(31), dict(zip(range(31), rng.permutation(31).tolist()))),
    }
    for G in out.values():
        G.faces = []
    return out
