#!/usr/bin/env python
"""bench.py -- DOF-updates/s of the gds stepping hot path on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat2d|hex_advdiff|ns_cavity|cml|wave3d] [--size S]
    python bench.py --impl reference ...     # the reference algorithm's CPU path (oracle port) on the host cores

A "step" is one fixed RK4 step (4 fused stage passes) -- or one map application for `cml` -- of the whole field.
Default workload: the heat law of BASELINE config 1 (`dydt = T.laplacian(y)`, static + time-varying Dirichlet) on
the 10^8-edge lattice `grid_2d_graph(7072, 7072)` the north-star roofline target is stated on.
`value` times K steps with the state resident in HBM (CUDA events on the library's stream, max over ranks);
`e2e` times the same steps through the public API with HOST state buffers: every step uploads the state from pinned
host memory, advances one step and reads the result back -- at N = 1 for an ensemble of independent host states through
`step_ensemble` (uploads, steps and read-backs of consecutive members overlapped; `e2e.serial` keeps the one-at-a-time
figure).  Inputs (>= 400 MB state + operators) exceed the 126 MB L2.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'dof_updates_per_sec'
UNIT = 'DOF-updates/s'


# ------------------------------------------------------------------------------------------------
# workloads: the same definition drives the CUDA product and the CPU oracle
# ------------------------------------------------------------------------------------------------
def heat2d_problem(api, n, native, world=1, rank=0, weighted=False, distribute=True):
    """BASELINE config 1 law on a grid (README.md:208-211 / examples/heat.py:56-63), RK4 h = 1e-2.  With `world` ranks
    the grid is (n*world) x n, cut into `world` slabs of n x n nodes (weak scaling: fixed work per GPU)."""
    h = 1e-2
    m = n * world
    if native:
        import gds_b200
        wts = None
        if weighted:   # same lattice with random edge weights: the general (weighted) operator layout
            wts = np.random.default_rng(5).uniform(0.5, 2.0, 2 * n * (n - 1) if world == 1 else 0)
        G = gds_b200.NativeGraph('grid_2d', m, n, weights=wts)
        if world > 1 and distribute:
            gds_b200.distribute(G, rank, world)
        T = api.node_gds(G)
        api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=h)
        i, j = np.divmod(T.global_ids, n)
        T.set_initial(y0=np.exp(-((i - m / 2) ** 2 + (j - n / 2) ** 2) / (n / 2)))
        jj = np.arange(n)
        idx = np.concatenate([jj, (m - 1) * n + jj])           # x[0] == 0 (time-varying), x[0] == m-1 (zero)
        T.set_constraints(dirichlet=gds_b200.BoundarySpec(
            idx, fun=lambda t: np.concatenate([np.sin(t + jj / 4) ** 2, np.zeros(n)])))
        n_own = T.ndim if T.partition is None else T.partition.n_own
    else:
        import cases
        G = cases.g_grid(n, n)
        T = api.node_gds(G)
        api.evolve(T, dydt=lambda t, y: T.laplacian(y), max_step=h)
        c = n / 2
        T.set_initial(y0=lambda x: math.exp(-((x[0] - c) ** 2 + (x[1] - c) ** 2) / (n / 2)))
        T.set_constraints(dirichlet=api.combine_bcs(
            lambda x: 0. if x[0] == n - 1 else None,
            lambda t, x: math.sin(t + x[1] / 4) ** 2 if x[0] == 0 else None))
        n_own = n * n
    V, E = n_own, 2 * n_own                                     # per rank; E ~ 2V on the 2-D grid
    if world == 1:
        E = 2 * n * (n - 1)
    m_op = 12 * (V + 2 * E) + 4 * (V + 1)                       # L0 as CSR (fp64 values, int32 indices)
    # bytes the kernel's own layout moves per step: 4 ELL columns of int32 indices (+ fp64 weights if weighted) per
    # stage, plus the RK4 vector traffic (x / y_n / acc reads, acc / x_next / y_new writes; stage 1 reads x = y_n once)
    # (a grid at most 32767 wide: the quad kernel reads 16-bit relative indices, 8 B per row and stage)
    k_op = ((8 if n <= 32767 else 16) + (32 if weighted else 0)) * V
    return dict(stepper=T, fields=[T], h=h, n_state=n_own, V=V, E=E, F=0,
                bytes_per_step=4 * m_op + 136 * V, bytes_per_launch=(4 * m_op + 136 * V) / 4.0,
                kernel_bytes_per_launch=(4 * k_op + 136 * V) / 4.0,
                dominant_class=0, kernel='lapq_kernel<unit weights>' if not weighted else 'lapq_kernel<weighted>',
                name=f'heat RK4 h=1e-2 on grid_2d_graph({m},{n}) (V={m * n}, E={2 * m * n - m - n}), Dirichlet static+time-varying'
                     + (', random edge weights' if weighted else '')
                     + (f', {world} slabs of {n}x{n} with per-stage halo exchange' if world > 1 else ''))


def _edges_inside(eu, ev, node_mask):
    """indices of the edges whose endpoints are both in the node set (= edges of G.subgraph(nodes))"""
    return np.nonzero(node_mask[eu] & node_mask[ev])[0]


def hex_advdiff_problem(api, n, native, world=1, rank=0, weighted=False, distribute=True):
    """BASELINE config 2: edge diffusion (Hodge-1 Laplacian, Dirichlet 0 on top/bottom boundary edges) coupled with node
    advection-diffusion (Neumann -0.1 on the top row) on hexagonal_lattice_graph(n, n), RK4 h = 1e-3 (tests/cases.py
    case_hex_advdiff_c2 is the same law at small size, checked against the reference)."""
    assert not weighted
    h, nu, kappa = 1e-3, 0.1, 0.05
    if native:
        import gds_b200
        # with `world` ranks: hexagonal_lattice_graph(n, n*world) cut into node-index ranges (vertical strips), 4 ghost
        # layers (hexagons: the edge <- face gather needs 3), edge and node state exchanged after every stage
        G = gds_b200.NativeGraph('hexagonal', n, n * world)
        if world > 1 and distribute:
            gds_b200.distribute(G, rank, world, hops=4)
        u, c = api.edge_gds(G), api.node_gds(G)
        hg = G.hostgraph(with_faces=True)
        pos, (eu, ev) = hg.pos(), hg.edges()
        api.evolve(u, dydt=lambda t, y: nu * u.laplacian(y), max_step=h)
        api.evolve(c, dydt=lambda t, y: -c.advect(u) + kappa * c.laplacian(y), max_step=h)
        u.set_initial(y0=0.1 * np.random.default_rng(0).standard_normal(hg.E))
        c.set_initial(y0=np.random.default_rng(1).random(hg.V))
        hh = math.sqrt(3) / 2
        lo, hi = pos[:, 1].min(), pos[:, 1].max()
        bnd = (pos[:, 1] < lo + 0.6 * hh) | (pos[:, 1] > hi - 0.6 * hh)
        u.set_constraints(dirichlet=gds_b200.BoundarySpec(_edges_inside(eu, ev, bnd), values=0.0))
        c.set_constraints(neumann=gds_b200.BoundarySpec(np.nonzero(pos[:, 1] > hi - 1e-9)[0], values=-0.1))
        sysobj = api.couple({'u': u, 'c': c})
        V, E = int(c.owned_mask.sum()), int(u.owned_mask.sum())      # this rank's share
        F = hg.F // world
    else:
        import cases
        import networkx as nx
        G = cases.g_hex(n, n)
        u, c = api.edge_gds(G), api.node_gds(G)
        api.evolve(u, dydt=lambda t, y: nu * u.laplacian(y), max_step=h)
        api.evolve(c, dydt=lambda t, y: -c.advect(u) + kappa * c.laplacian(y), max_step=h)
        u0 = 0.1 * np.random.default_rng(0).standard_normal(u.ndim)
        c0 = np.random.default_rng(1).random(c.ndim)
        u.set_initial(y0=lambda e: u0[u.X[e]])
        c.set_initial(y0=lambda x: c0[c.X[x]])
        u.set_constraints(dirichlet=api.zero_edge_bc(cases.boundary_edges_hex_top_bottom(G)))
        pos = nx.get_node_attributes(G, 'pos')
        ytop = max(p[1] for p in pos.values())
        c.set_constraints(neumann=lambda x: -0.1 if pos[x][1] > ytop - 1e-9 else None)
        sysobj = api.couple({'u': u, 'c': c})
        V, E, F = c.ndim, u.ndim, len(u.faces)
    # edge law 64E + 20V + 44F, node law 72E + 28V per RHS; RK4 vectors 136 B per state component
    per_step = 4 * ((64 * E + 20 * V + 44 * F) + (72 * E + 28 * V)) + 136 * (E + V)
    # what the kernels' own layout moves (unit weights: no weight streams; int32 indices, 3 / 6 / 2 slots per node / face /
    # edge): per stage node divergence + face curl + edge RHS + node RHS with their RK vectors, once per step the advection term
    own = 4 * (104 * V + 80 * E + 40 * F) + (40 * V + 8 * E)
    return dict(stepper=sysobj.stepper, fields=[u, c], h=h, n_state=E + V, V=V, E=E, F=F, bytes_per_step=per_step,
                kernel_bytes_per_step=own,
                bytes_per_launch=per_step / 4.0, kernel_bytes_per_launch=own / 4.0, dominant_class=-1,
                kernel='all passes of one RK stage (terms_kernel: node divergence, face curl, edge RHS; lapT_kernel: '
                       'node RHS)',
                name=f'edge diffusion + node advection-diffusion, RK4 h=1e-3, hexagonal_lattice_graph({n},{n * world}) '
                     + (f'in {world} deep partitions (4 ghost layers), per rank ' if world > 1 else '')
                     + f'(V={V}, E={E}, F={F}), Dirichlet edges / Neumann nodes')


def ns_cavity_problem(api, n, native, world=1, rank=0, weighted=False):
    """BASELINE config 3: Navier-Stokes velocity law (examples/fluid.py:33-34) on triangular_lattice_graph(n, n), lid
    const_edge_bc (time-varying form), no-slip walls (examples/lid_driven_cavity.py:17-27), pressure held as a given node
    vector; accepted-state operators as in the reference law, RK4 h = 1e-4 (tests/cases.py case_ns_cavity_c3)."""
    assert world == 1 and not weighted
    h, rho, visc = 1e-4, 0.1, 200.0
    if native:
        import gds_b200
        G = gds_b200.NativeGraph('triangular', n, n)
        p, u = api.node_gds(G), api.edge_gds(G)
        hg = G.hostgraph(with_faces=True)
        pos, (eu, ev) = hg.pos(), hg.edges()
        api.evolve_nil(p)
        p.set_initial(y0=0.01 * np.random.default_rng(3).standard_normal(p.ndim))
        api.evolve(u, dydt=lambda t, y: -u.advect() - p.grad() / rho + u.laplacian() * visc / rho, max_step=h)
        u.set_initial(y0=0.05 * np.random.default_rng(4).standard_normal(u.ndim))
        x0, x1, y0_, y1 = pos[:, 0].min(), pos[:, 0].max(), pos[:, 1].min(), pos[:, 1].max()
        lid = _edges_inside(eu, ev, pos[:, 1] > y1 - 1e-9)
        walls = np.unique(np.concatenate([_edges_inside(eu, ev, pos[:, 1] < y0_ + 1e-9),
                                          _edges_inside(eu, ev, pos[:, 0] < x0 + 0.51),
                                          _edges_inside(eu, ev, pos[:, 0] > x1 - 0.51)]))
        lid = np.setdiff1d(lid, walls)                     # static (wall) conditions win (boundary.py:21-41)
        idx = np.concatenate([walls, lid])
        vals = np.concatenate([np.zeros(walls.size), np.full(lid.size, 10.0)])
        u.set_constraints(dirichlet=gds_b200.BoundarySpec(idx, fun=lambda t: vals))
        sysobj = api.couple({'u': u, 'p': p})
        V, E, F = hg.V, hg.E, hg.F
    else:
        import cases
        G = cases.g_tri(n, n)
        p, u = api.node_gds(G), api.edge_gds(G)
        p0 = 0.01 * np.random.default_rng(3).standard_normal(p.ndim)
        api.evolve_nil(p)
        p.set_initial(y0=lambda x: p0[p.X[x]])
        api.evolve(u, dydt=lambda t, y: -u.advect() - p.grad() / rho + u.laplacian() * visc / rho, max_step=h)
        u0 = 0.05 * np.random.default_rng(4).standard_normal(u.ndim)
        u.set_initial(y0=lambda e: u0[u.X[e]])
        lid, walls = cases.tri_boundaries(G)
        lid_bc = api.const_edge_bc(lid, 10.0)
        u.set_constraints(dirichlet=api.combine_bcs(api.zero_edge_bc(walls), lambda t, e: lid_bc(e)))
        sysobj = api.couple({'u': u, 'p': p})
        V, E, F = p.ndim, u.ndim, len(u.faces)
    # advect 48E + 76V, grad 16E + 8V, Hodge-1 Laplacian 64E + 20V + 32F per RHS
    per_step = 4 * (128 * E + 104 * V + 32 * F) + 136 * E
    # the law reads accepted state only (fds.py:377-378): its operators are evaluated once per step and
    # the four stage combines happen in registers (collapsed RK4)
    hoisted = (128 * E + 104 * V + 32 * F) + 16 * E      # + read y_n, write y_{n+1}
    # the kernels' own layout, collapsed step: aggregates, divergence, curl, edge law (unit weights, int32 indices)
    own = 144 * V + 64 * E + 28 * F
    return dict(stepper=sysobj.stepper, fields=[u], h=h, n_state=E, V=V, E=E, F=F, bytes_per_step=per_step,
                bytes_per_step_hoisted=hoisted, kernel_bytes_per_step=own,
                bytes_per_launch=per_step / 4.0, kernel_bytes_per_launch=own, dominant_class=-1,
                kernel='step level: the law reads accepted state only, so the step is collapsed -- node aggregates, '
                       'divergence, curl and the edge law run ONCE per step, the four RK4 stage combines in registers',
                name=f'Navier-Stokes edge law (advect + grad p + Hodge-1 Laplacian), RK4 h=1e-4, '
                     f'triangular_lattice_graph({n},{n}) (V={V}, E={E}, F={F}), lid + no-slip Dirichlet')


def cml_problem(api, n, native, world=1, rank=0, weighted=False, distribute=True):
    """BASELINE config 4: coupled map lattice on random_geometric_graph(n, sqrt(8/(pi n)), seed=42), map mode.  With
    `world` ranks the graph has n * world nodes (--total: n) and is cut into spatially compact parts along a space-filling
    curve over the node positions; one halo exchange per map application."""
    assert not weighted
    nn = n if (TOTAL[0] or not native) else n * world
    r = math.sqrt(8 / (math.pi * nn))
    eps = 0.3
    if native:
        import gds_b200
        G = gds_b200.NativeGraph('random_geometric', nn, r, 42)
        if world > 1 and distribute:
            gds_b200.distribute(G, rank, world, align=32)
    else:
        import cases
        G = cases.g_rgg(n, r, 42)
    x = api.node_gds(G)
    api.evolve_map(x, map_fun=lambda t, y: (1 - 1.7 * y ** 2) + (eps / 8.0) * x.laplacian(1 - 1.7 * y ** 2), dt=1.0)
    y0 = np.random.default_rng(2).uniform(-1, 1, nn)
    x.set_initial(y0=y0 if native else (lambda v: y0[v]))
    part = getattr(x, 'partition', None) if native else None
    V = nn if part is None else part.n_own
    E = (x._low.E if part is None else int(x._low.E * V / max(1, x._low.V))) if native else G.number_of_edges()
    per_step = 12 * (V + 2 * E) + 4 * V + 16 * V
    return dict(stepper=x, fields=[x], h=1.0, n_state=V, V=V, E=E, F=0, bytes_per_step=per_step,
                kernel_bytes_per_step=4 * 2 * E + 24 * V,      # int32 indices, gather of f(y) once, y and f(y) written
                bytes_per_launch=per_step, kernel_bytes_per_launch=(4 * 2 * E + 8 * 4) * 1.0 + 24 * V,
                dominant_class=0, kernel='lapT_kernel<unit weights, map> (+ pointwise pass f(y))',
                name=f'coupled map lattice, random_geometric_graph({nn}) (E={E if part is None else "~" + str(E * world)}), map mode'
                     + (f', {world} parts along a space-filling curve, halo exchange per application' if part is not None else ''))


def wave3d_problem(api, n, native, world=1, rank=0, weighted=False, distribute=True):
    """BASELINE config 5 law: wave equation, order 2, accepted-state Laplacian (examples/wave.py:14) on grid_graph([n]*3)"""
    assert not weighted
    h = 0.1
    if native:
        import gds_b200
        # weak scaling: n x n x (n * world) cut into `world` slabs of n^3; --total: the n^3 grid itself cut into slabs
        # (strong scaling -- BASELINE config 5 is the 693^3 grid, 10^9 edges, on 8 GPUs)
        nz = n if TOTAL[0] else n * world
        G = gds_b200.NativeGraph('grid', n, n, nz)
        if world > 1 and distribute:
            # slabs along the first label axis by default; --blocks pa,pb,pc: blocks (2 x 2 x 2 for config 5)
            gds_b200.distribute(G, rank, world, blocks=BLOCKS[0])
    else:
        import cases
        G = cases.g_grid3((n, n, n))
    a = api.node_gds(G)
    api.evolve(a, dydt=lambda t, y: 1.0 * a.laplacian(), order=2, max_step=h)
    if native:
        gid = a.global_ids
        A, B, C = nz, n, n
        ia, ib, ic = gid // (B * C), (gid // C) % B, gid % C
        a.set_initial(y0=np.exp(-((ia - (A - 1) / 2) ** 2 + (ib - (B - 1) / 2) ** 2 + (ic - (C - 1) / 2) ** 2) / 200.0))
        n_own = a.ndim if a.partition is None else a.partition.n_own
    else:
        ctr = (n - 1) / 2
        a.set_initial(y0=lambda x: math.exp(-sum((xi - ctr) ** 2 for xi in x) / 200.0))
        n_own = a.ndim
    V, E = n_own, 3 * n_own
    per_step = 4 * (12 * (V + 2 * E) + 4 * V) + 136 * 2 * V        # Laplacian at every stage
    # `a.laplacian()` reads the accepted state: one Laplacian per step (collapsed RK4)
    hoisted = (12 * (V + 2 * E) + 4 * V) + 24 * V + 32 * V      # Laplacian pass: read v_n, write k_v, v_{n+1}; shift pass: read k_v, v_n, x_n, write x_{n+1}
    own = (4 * 6 + 32) * V        # one pass per step: 6 int32 indices, read x and v_n, write x_{n+1} and v_{n+1}
    return dict(stepper=a, fields=[a], h=h, n_state=2 * n_own, V=V, E=E, F=0, bytes_per_step=per_step,
                bytes_per_step_hoisted=hoisted, kernel_bytes_per_step=own,
                bytes_per_launch=per_step / 4.0, kernel_bytes_per_launch=own, dominant_class=-1,
                kernel='step level: collapsed RK4 -- ONE Laplacian pass per step that advances the derivative block and, in '
                       'the same epilogue, the value block',
                name=f'wave equation order 2, RK4 h=0.1, grid_graph([{n},{n},{nz if native else n}]) '
                     f'(V={n * n * (nz if native else n)}, E={3 * n * n * (nz if native else n) - n * n - 2 * n * (nz if native else n)})'
                     + (f', {world} ' + ('blocks ' + 'x'.join(map(str, BLOCKS[0])) if BLOCKS[0] else 'slabs') + ' with halo exchange'
                        if world > 1 else ''))


# name -> (builder, default GPU size per rank, CPU sample size)
WORKLOADS = {'heat2d': (heat2d_problem, 7072, 1024), 'hex_advdiff': (hex_advdiff_problem, 2000, 40),
             'ns_cavity': (ns_cavity_problem, 4000, 40),
             'cml': (cml_problem, 10_000_000, 100_000), 'wave3d': (wave3d_problem, 400, 48)}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.device}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
# sizes the UNMODIFIED reference can hold (it densifies the |V| x |E| incidence matrix, gds/gds.py:145,265-267, and
# assembles edge operators with O(E^2) Python loops): heat2d 100 is BASELINE config 1 exactly
REF_SIZES = {'heat2d': 100, 'hex_advdiff': 12, 'ns_cavity': 12, 'cml': 3000, 'wave3d': 16}


def reference_available():
    try:
        from oracle.ref_harness import reference_root
        return reference_root() is not None
    except Exception:
        return False


def cpu_run(workload, size, steps, warmup, kind='port'):
    """kind 'reference': the unmodified reference (oracle/_ref or /root/reference) driven by a fixed-step RK4 solver
    object; 'port': the oracle (numpy / scipy restatement)"""
    if kind == 'reference':
        import warnings
        warnings.simplefilter('ignore')
        from oracle.ref_api import make_ref_api
        api = make_ref_api()
    else:
        from adaptors import oracle_api
        api = oracle_api()
    builder = WORKLOADS[workload][0]
    prob = builder(api, size, native=False)
    st, h = prob['stepper'], prob['h']
    for _ in range(warmup):
        st.step(h)
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(h)
    el = time.perf_counter() - t0
    return prob, el


def _replica_worker(workload, size, kind, steps, barrier, q):
    """one independent replica of the CPU arm (its own process, one core)"""
    try:
        import warnings
        warnings.simplefilter('ignore')
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)                      # the reference prints ("Faces: n") on stdout
        prob, _ = cpu_run(workload, size, steps=1, warmup=1, kind=kind)
        st, h = prob['stepper'], prob['h']
        barrier.wait(timeout=90)
        t0 = time.perf_counter()
        for _ in range(steps):
            st.step(h)
        q.put((prob['n_state'] * steps, t0, time.perf_counter()))
    except Exception as exc:                     # a replica that fails must not hang the others at the barrier
        try:
            barrier.abort()
        except Exception:
            pass
        q.put(('error', f'{type(exc).__name__}: {exc}'[:200]))


def cpu_all_cores(workload, size, kind, steps):
    """The reference's stepping path (scipy CSR SpMV + numpy) is single-threaded, so one process uses one core.  For a
    figure that uses the whole host, one independent replica of the same problem per core runs concurrently (the CPU
    counterpart of stepping an ensemble): aggregate DOF-updates over the span from the first start to the last finish.
    Replicas are capped by the memory the reference needs (it densifies the incidence matrix: ~2.5 GB at grid 100x100)."""
    import multiprocessing as mp
    n = os.cpu_count() or 1
    try:
        import psutil
        per = 3.0e9 if kind == 'reference' else 1.5e9
        n = max(1, min(n, int(psutil.virtual_memory().available * 0.6 / per)))
    except Exception:
        n = min(n, 8)
    if n < 2:
        return None
    ctx = mp.get_context('spawn')      # fresh interpreters: the GPU arm's process holds a CUDA context that must not be forked
    barrier, q = ctx.Barrier(n), ctx.Queue()
    procs = [ctx.Process(target=_replica_worker, args=(workload, size, kind, steps, barrier, q)) for _ in range(n)]
    for pr in procs:
        pr.start()
    res, deadline = [], time.perf_counter() + 120.0        # a bounded extra: never more than two minutes of the run
    try:
        for _ in procs:
            res.append(q.get(timeout=max(1.0, deadline - time.perf_counter())))
    except Exception:
        pass
    for pr in procs:
        pr.join(timeout=30)
        if pr.is_alive():
            pr.terminate()
    ok = [r for r in res if r[0] != 'error']
    if len(ok) != n:
        return {'error': (res and [r for r in res if r[0] == 'error'][:1]) or 'a replica did not finish', 'processes': n}
    span = max(r[2] for r in ok) - min(r[1] for r in ok)
    return {'value': sum(r[0] for r in ok) / span, 'unit': UNIT, 'cores': n, 'processes': n, 'steps_per_process': steps,
            'note': 'one independent replica of the same problem per core, run concurrently (the stepping path itself is '
                    'single-threaded); aggregate over first start .. last finish'}


def _cpu_sample(workload, size, kind, budget_s):
    prob, el1 = cpu_run(workload, size, steps=3, warmup=1, kind=kind)
    steps = int(max(3, min(2000, budget_s / max(el1 / 3, 1e-6))))
    st, h = prob['stepper'], prob['h']
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(h)
    el = time.perf_counter() - t0
    what = ('the unmodified reference (asrvsn/gds, oracle/_ref) with an RK4 solver object swapped in; its stepping path is '
            'scipy CSR SpMV + numpy, single-threaded' if kind == 'reference'
            else 'oracle/gds_oracle.py, scipy CSR, single-threaded')
    return {'value': prob['n_state'] * steps / el, 'unit': UNIT, 'cores': 1, 'kind': kind, 'steps_per_second': steps / el,
            'sample': f'{prob["name"]}; {steps} steps in {el:.2f} s ({what}; {os.cpu_count()} host cores present)'}


def cpu_baseline(workload, size, budget_s=12.0):
    """bounded samples of the same law on the host: the reference itself at the size it can hold (BASELINE config 1 for
    the heat law) when oracle/_ref is present, and the oracle port at a larger size"""
    port = _cpu_sample(workload, size, 'port', budget_s)
    if not reference_available():
        return port
    ref = _cpu_sample(workload, REF_SIZES[workload], 'reference', budget_s)
    ref['port'] = port
    try:
        steps = int(max(20, min(4000, 4.0 * ref['steps_per_second'])))        # about 4 s per replica
        ref['all_cores'] = cpu_all_cores(workload, REF_SIZES[workload], 'reference', steps)
    except Exception as exc:
        ref['all_cores'] = {'error': f'{type(exc).__name__}: {exc}'[:200]}
    return ref


_REAL_STDOUT = [None]


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout at
    the first collective), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a saved
    copy of the real stdout."""
    sys.stdout.flush()
    _REAL_STDOUT[0] = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    fd = _REAL_STDOUT[0] if _REAL_STDOUT[0] is not None else 1
    os.write(fd, (line + '\n').encode())


def reference_arm(args):
    """the reference's own CPU implementation of the path on the host cores: the UNMODIFIED reference where oracle/_ref
    (or /root/reference) is present -- at the size it can hold -- else the oracle port"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    workload = args.workload
    kind = 'reference' if reference_available() and not args.cpu_size else 'port'
    size = REF_SIZES[workload] if kind == 'reference' else (args.cpu_size or WORKLOADS[workload][2])
    prob, el = cpu_run(workload, size, args.steps, args.warmup, kind=kind)
    v = prob['n_state'] * args.steps / el
    what = ('the unmodified reference (asrvsn/gds, imported from oracle/_ref) with an RK4 solver object swapped in '
            '(gds/ode/midpoint.py protocol)' if kind == 'reference'
            else 'oracle/gds_oracle.py (numpy/scipy restatement of the reference)')
    cb = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': kind,
          'sample': f'{prob["name"]}; {what}; the stepping path is scipy CSR SpMV + numpy, which is single-threaded '
                    f'({os.cpu_count()} host cores present)'}
    if kind == 'reference' and not args.no_cpu:
        # the line's value stays what the reference does (one thread); the whole-host figure beside it
        try:
            sps = args.steps / el
            cb['all_cores'] = cpu_all_cores(workload, size, kind, int(max(20, min(4000, 4.0 * sps))))
        except Exception as exc:
            cb['all_cores'] = {'error': f'{type(exc).__name__}: {exc}'[:200]}
    emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': prob['name'], 'note': 'bounded CPU sample of the GPU arm\'s law at the size the reference '
                                                     'can hold (it densifies the incidence matrix)'},
        'cpu_baseline': cb,
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_steps(eng, ctx, K, h, barrier, device, world, dist, torch, est_ms=2.0):
    """K steps with the state resident in HBM: CUDA events on the library's stream, max over ranks.  The clock sampler
    (100 ms period) needs load for longer than the timed region, so the same steps first run untimed for about 0.3 s
    and the timed K steps follow without a gap."""
    n_pre = 1 if NO_PREROLL[0] else max(1, min(2000, int(0.3 / max(1e-6, est_ms * 1e-3))))
    if world > 1:
        # every rank must enqueue the SAME number of steps: the fused halo exchange pairs them up one by one, and
        # `est_ms` is a rank-local measurement
        tt = torch.tensor([n_pre], dtype=torch.int64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        n_pre = int(tt.item())
    with ClockSampler(device) as clk:
        eng.run(eng.plan_steps(n_pre, h))
        plan = eng.plan_steps(K, h)
        launches0, _ = ctx.counters()
        barrier()
        ctx.timer_start()
        eng.run(plan)
        ms = ctx.timer_stop()
        barrier()
        launches1, _ = ctx.counters()
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms, int(launches1 - launches0), clk.summary()


def verify_partitioned(eng, builder, args, world, rank, dist, torch):
    """Correctness of the partitioned run, checked in the bench itself (the driver only sees this line at N > 1):
    (1) halo_check -- after the timed steps every rank fetches, over NCCL send/recv, the rows its neighbours own and
        compares them bit for bit with the ghost rows the fused peer-to-peer exchange left on it (full bench size);
    (2) slab_check -- the same law on a small lattice (slabs still > 32767 rows, so ghost rows sit behind the 16-bit
        index escape), 6 steps partitioned over all ranks vs. the whole lattice on rank 0's GPU alone: owned rows of
        every rank bit-identical;
    (3) the error flag and exchange count of the fused exchange."""
    from adaptors import product_api
    out = {'transport': 'p2p' if getattr(eng, 'p2p', False) else 'host', 'overlap_split': getattr(eng, 'overlap', None)}
    eng.sync()
    bad = torch.tensor([eng.halo_check()], dtype=torch.int64, device='cuda')
    dist.all_reduce(bad)
    out['halo_check'] = 'bit-identical' if int(bad.item()) == 0 else f'{int(bad.item())} ghost values differ from their owners'
    if getattr(eng, 'p2p', False):
        ex, err = eng.p2p_status()
        st = torch.tensor([ex, err], dtype=torch.int64, device='cuda')
        every = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(every, st)
        out['p2p'] = {'exchanges_per_rank': [int(x[0].item()) for x in every], 'error': max(int(x[1].item()) for x in every)}
    small = {'heat2d': 512, 'wave3d': 48, 'hex_advdiff': 24, 'cml': 200_000}.get(args.workload)
    slab_ok = True
    if small is not None and not args.no_slab_check:
        steps = 6
        prob = builder(product_api(), small, native=True, world=world, rank=rank)
        e2 = prob['stepper']._ensure_engine()
        e2.run(e2.plan_steps(steps, prob['h']))
        e2.sync()
        mine = [(np.asarray(fl.global_ids)[fl.owned_mask], np.asarray(fl.y_owned)) for fl in prob['fields']]
        bad2 = torch.tensor([e2.halo_check()], dtype=torch.int64, device='cuda')
        dist.all_reduce(bad2)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        verdict = [1]
        if rank == 0:
            one = builder(product_api(), small, native=True, world=world, rank=0, distribute=False)
            e1 = one['stepper']._ensure_engine()
            e1.run(e1.plan_steps(steps, one['h']))
            e1.sync()
            for k, fl in enumerate(one['fields']):
                want = np.asarray(fl.y)
                got = np.full_like(want, np.nan)
                for r in range(world):
                    gid, val = gathered[r][k]
                    got[gid] = val
                if not np.array_equal(got, want):
                    verdict[0] = 0
        dist.broadcast_object_list(verdict, src=0)
        slab_ok = bool(verdict[0]) and int(bad2.item()) == 0
        out['slab_check'] = (('bit-identical' if slab_ok else 'MISMATCH') + f' ({prob["name"]}; {steps} steps, {world} ranks vs one GPU;'
                             f' overlap split {getattr(e2, "overlap", None)}, ghost rows vs owners: {int(bad2.item())} differ)')
    out['ok'] = int(bad.item()) == 0 and slab_ok and out.get('p2p', {}).get('error', 0) == 0
    return out


def roofline_of(prob, eng, K, h, ms_step):
    """per-launch time of the dominant kernel: a CUDA-event pair around every launch, eager replay of the same steps"""
    cls_ms, cls_n = eng.run_timed(eng.plan_steps(min(K, 64), h))
    dom = prob['dominant_class']
    if dom < 0:      # several kernels per stage: one "launch" = all passes of one stage
        n_st = getattr(eng, 'n_stages', 4 if str(eng.scheme).endswith('rk4') else 1)
        launch_ms = float(cls_ms.sum()) / (min(K, 64) * n_st)
        cls_ms = np.array([cls_ms.sum(), 0.0])
        dom = 0
    else:
        launch_ms = cls_ms[dom] / max(1, cls_n[dom])
    peak, peak_src = measured_peak()
    achieved = prob['bytes_per_launch'] / (launch_ms * 1e-3) / 1e9
    kern = prob['kernel_bytes_per_launch'] / (launch_ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': None, 'peak_source': peak_src, 'kernel': prob['kernel'],
            'algorithmic_bytes_per_launch': prob['bytes_per_launch'], 'launch_ms': launch_ms,
            'kernel_bytes_per_launch': prob['kernel_bytes_per_launch'], 'achieved_kernel_bytes': kern,
            'frac_kernel_bytes': kern / peak, 'kernel_share_of_step': float(cls_ms[dom] / cls_ms.sum()),
            'timing': 'CUDA-event pair around every launch, eager replay of the same steps',
            'step_level_gbs': prob['bytes_per_step'] / (ms_step * 1e-3) / 1e9,
            'note': 'achieved/frac use the ALGORITHMIC bytes of SURVEY.md 8d (general weighted CSR); on a unit-weight '
                    'graph the kernel stores no weights and moves kernel_bytes_per_launch, see frac_kernel_bytes'}


def other_workloads(args, ctx, barrier, local, dist, torch, skip):
    """The other BASELINE configurations at their full sizes, one after another on the same GPU (N = 1): K steps with the
    state resident in HBM, CUDA events, clocks sampled -- same protocol as the headline.  frac_algorithmic uses SURVEY.md
    8d's bytes (operators counted at every stage, general weighted CSR) and may exceed 1; laws that read accepted state only
    are evaluated once per step by the engine (collapsed RK4): frac_hoisted counts SURVEY's bytes with the operators once;
    frac_kernel_bytes counts what the kernels' own layout has to move (unit weights: no weight streams) -- the honest
    bandwidth fraction by design; frac_traffic uses the DRAM bytes of one step from the committed ncu capture
    (profiles/traffic.json), when there is one."""
    import gc
    from adaptors import product_api
    peak, _ = measured_peak()
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            traffic = json.load(fh)
    except Exception:
        traffic = {}
    out = []
    for name in ('hex_advdiff', 'ns_cavity', 'cml', 'wave3d'):
        if name == skip:
            continue
        builder, size, _ = WORKLOADS[name]
        t0 = time.perf_counter()
        try:
            prob = builder(product_api(), size, native=True)
            eng = prob['stepper']._ensure_engine()
            h, K = prob['h'], max(3, min(args.steps, 10))
            barrier()
            ctx.timer_start()
            eng.run(eng.plan_steps(3, h))
            est = ctx.timer_stop() / 3
            t_build = time.perf_counter() - t0
            ms, launches, clocks = time_steps(eng, ctx, K, h, barrier, local, 1, dist, torch, est_ms=est)
            ms_step = ms / K
            rec = {'name': prob['name'], 'workload': name, 'size': size, 'steps': K, 'ms_per_step': ms_step,
                   'value': prob['n_state'] * K / (ms * 1e-3), 'unit': UNIT,
                   'algorithmic_bytes_per_step': prob['bytes_per_step'],
                   'frac_algorithmic': prob['bytes_per_step'] / (ms_step * 1e-3) / 1e9 / peak,
                   'frac_hoisted': (prob['bytes_per_step_hoisted'] / (ms_step * 1e-3) / 1e9 / peak
                                    if 'bytes_per_step_hoisted' in prob else None),
                   'kernel_bytes_per_step': prob.get('kernel_bytes_per_step'),
                   'frac_kernel_bytes': (prob['kernel_bytes_per_step'] / (ms_step * 1e-3) / 1e9 / peak
                                         if prob.get('kernel_bytes_per_step') else None),
                   'kernels_per_step': eng.kernels_per_step, 'gpu_launches': launches, 'build_seconds': t_build,
                   'clocks': clocks, 'kernel': prob['kernel']}
            tr = traffic.get(name, {}).get('per_step', {}).get(str(size))
            rec['traffic_bytes_per_step'] = tr
            rec['frac_traffic'] = tr / (ms_step * 1e-3) / 1e9 / peak if tr else None
            out.append(rec)
            del eng
            prob.clear()
        except Exception as exc:          # one workload failing must not cost the headline line
            out.append({'workload': name, 'error': f'{type(exc).__name__}: {exc}'[:300]})
        gc.collect()
    if skip != 'ns_cavity' and not args.no_pressure:
        try:
            out.append(pressure_workload(ctx, WORKLOADS['ns_cavity'][1], max(3, min(args.steps, 5))))
        except Exception as exc:
            out.append({'workload': 'ns_pressure', 'error': f'{type(exc).__name__}: {exc}'[:300]})
        gc.collect()
    return out


def pressure_workload(ctx, n, steps):
    """BASELINE config 3 with the pressure SOLVED every step (examples/fluid.py:14-38 `navier_stokes`, lid-driven cavity of
    examples/lid_driven_cavity.py:56-67: pressure reference at one node): the `lhs=` pressure field is a preconditioned
    conjugate-gradient solve on the device (gds_pcg_solve, smoothed-aggregation V-cycle) before every internal step
    (fds.py:505-507), the velocity law the collapsed RK4 step of ns_cavity.  Wall clock around the public step() calls --
    every solve reads its residual back, so the calls are synchronous."""
    import gds_b200
    from adaptors import product_api
    api = product_api()
    h, rho, visc, min_step = 1e-4, 0.1, 200.0, 1e-3
    t0 = time.perf_counter()
    G = gds_b200.NativeGraph('triangular', n, n)
    p, u = api.node_gds(G), api.edge_gds(G)
    hg = G.hostgraph(with_faces=True)
    pos, (eu, ev) = hg.pos(), hg.edges()

    def pressure_f(t, y):
        dt = max(min_step, u.dt)
        lhs = u.div(u.y / dt - u.advect() + u.laplacian() * visc / rho)
        lhs -= p.laplacian(y) / rho
        return lhs
    api.evolve_lhs(p, pressure_f)
    api.evolve(u, dydt=lambda t, y: -u.advect() - p.grad() / rho + u.laplacian() * visc / rho, max_step=h)
    u.set_initial(y0=0.05 * np.random.default_rng(4).standard_normal(u.ndim))
    x0, x1, y0_, y1 = pos[:, 0].min(), pos[:, 0].max(), pos[:, 1].min(), pos[:, 1].max()
    lid = _edges_inside(eu, ev, pos[:, 1] > y1 - 1e-9)
    walls = np.unique(np.concatenate([_edges_inside(eu, ev, pos[:, 1] < y0_ + 1e-9),
                                      _edges_inside(eu, ev, pos[:, 0] < x0 + 0.51),
                                      _edges_inside(eu, ev, pos[:, 0] > x1 - 0.51)]))
    lid = np.setdiff1d(lid, walls)
    u.set_constraints(dirichlet=gds_b200.BoundarySpec(np.concatenate([walls, lid]),
                                                      values=np.concatenate([np.zeros(walls.size), np.full(lid.size, 10.0)])))
    p.set_constraints(dirichlet=gds_b200.BoundarySpec(np.array([0]), values=0.0))      # pressure reference
    sysobj = api.couple({'u': u, 'p': p})
    sysobj.step(h)                                   # builds the hierarchy, the pass lists and the captured step
    ctx.sync()
    t_build = time.perf_counter() - t0
    sysobj.step(h)
    ctx.sync()
    its, solve_s = [], []
    t0 = time.perf_counter()
    for _ in range(steps):
        sysobj.step(h)
        its.append(p.last_solve['iterations'])
        solve_s.append(p.last_solve['solve_seconds'])
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    from gds_b200.multigrid import Multigrid
    info = Multigrid.for_graph(p._low, p.dirichlet_indices).info()
    return {'name': f'Navier-Stokes with the pressure solved every step (lhs= field, multigrid-preconditioned CG) + velocity '
                    f'RK4 h=1e-4, triangular_lattice_graph({n},{n}) (V={hg.V}, E={hg.E}, F={hg.F}), lid-driven cavity',
            'workload': 'ns_pressure', 'size': n, 'steps': steps, 'ms_per_step': ms, 'value': hg.E * steps / (ms * steps * 1e-3),
            'unit': UNIT, 'solve_iterations': its, 'solve_ms': [1e3 * x for x in solve_s],
            'relative_residual': p.last_solve['relative_residual'], 'hierarchy': info, 'build_seconds': t_build,
            'timing': 'wall clock around the public step() calls (each solve synchronises)'}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference for the CPU path)')
    torch.cuda.set_device(local)
    os.environ['GDS_B200_DEVICE'] = str(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import gds_b200  # noqa: F401
    from adaptors import product_api
    builder, size_default, cpu_size = WORKLOADS[args.workload]
    size = args.size or size_default
    t_build = time.perf_counter()
    prob = builder(product_api(), size, native=True, world=world, rank=rank, weighted=args.weighted)
    WORKLOAD_NAME[0] = prob['name']
    st, h, K, W = prob['stepper'], prob['h'], args.steps, args.warmup
    eng = st._ensure_engine()
    SCHEME[0] = 'map' if eng.scheme == 'map' else str(eng.scheme).split('.')[-1]
    HALO[0] = ('peer-to-peer stores over NVLink fused into the captured step (CUDA IPC, no NCCL on the data path)'
               if getattr(eng, 'p2p', False) else 'pack kernel, NCCL send/recv, unpack kernel, driven by the host')
    ctx = eng.ctx
    t_build = time.perf_counter() - t_build

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- kernel-resident throughput: K steps, state in HBM ----------------------------------------------------
    barrier()
    ctx.timer_start()
    eng.run(eng.plan_steps(W, h))
    est = ctx.timer_stop() / W
    ms, launches, clocks = time_steps(eng, ctx, K, h, barrier, local, world, dist, torch, est_ms=est)
    n_total = prob['n_state'] * world
    if world > 1:       # slabs may differ by a lattice plane: count what the ranks really advance
        tt = torch.tensor([prob['n_state']], dtype=torch.int64, device='cuda')
        dist.all_reduce(tt)
        n_total = int(tt.item())
    value = n_total * K / (ms * 1e-3)
    multi = None
    if world > 1:
        multi = verify_partitioned(eng, builder, args, world, rank, dist, torch)
    roof = roofline_of(prob, eng, K, h, ms / K)
    if world > 1:   # GPUs of one box differ by several per cent; the step runs at the pace of the slowest
        mine = torch.tensor([roof['launch_ms']], dtype=torch.float64, device='cuda')
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        roof['launch_ms_per_rank'] = [float(x.item()) for x in every]
    try:   # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            roof['traffic'] = json.load(fh).get(args.workload, {}).get(str(size) + ('_heat_weighted' if args.weighted else ''))
    except Exception:
        pass

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    fields = prob['fields']
    # host buffers per field: the rows this rank owns (the whole local vector where owned rows are not a prefix)
    sizes = [fl.ndim if (getattr(fl, 'partition', None) is None or hasattr(fl.partition, 'gid_dom')) else fl.partition.n_own
             for fl in fields]
    n = sum(s_ * fl._order() for s_, fl in zip(sizes, fields))
    host_in = [torch.empty(s_, dtype=torch.float64).pin_memory().numpy() for s_ in sizes]
    host_out = [torch.empty(s_, dtype=torch.float64).pin_memory().numpy() for s_ in sizes]
    for fl, b in zip(fields, host_in):
        eng.read_into(fl, b)
    Ke = max(3, min(K, 10))

    def one_step(src, dst):
        for fl, b in zip(fields, src):
            eng.write(fl, b)           # H2D of the step's input state (pinned)
        st.step(h)                     # public API call
        for fl, b in zip(fields, dst):
            eng.read_into(fl, b)       # D2H of the step's result (synchronises)
    for _ in range(2):
        one_step(host_in, host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        one_step(host_in, host_out)
        host_in, host_out = host_out, host_in
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {'value': n_total * Ke / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': 8 * sum(sizes), 'd2h_bytes_per_step': 8 * sum(sizes),
           'steps': Ke, 'note': 'per step: pinned-host state upload, one public step() call, state read-back; wall clock'}
    if world == 1 and all(fl._order() == 1 for fl in fields) and hasattr(st, 'step_ensemble') and eng.scheme != 'map':
        # The same per-step work -- a host state goes up, one step, the new state comes back -- for an ENSEMBLE of
        # independent host-resident states through the public step_ensemble call: uploads, steps and read-backs of
        # consecutive members overlap (both PCIe directions busy at once).  Two distinct input states alternate; every
        # member's 8 * n bytes cross PCIe in each direction inside the timed region.
        try:
            serial = dict(e2e)
            in2 = [torch.empty(s_, dtype=torch.float64).pin_memory().numpy() for s_ in sizes]
            out2 = [torch.empty(s_, dtype=torch.float64).pin_memory().numpy() for s_ in sizes]
            for a, b in zip(in2, host_out):
                a[:] = b                      # a second, different state (one step later)
            from gds_b200.fields import fds as _fds
            single = isinstance(st, _fds)
            wrap = (lambda grp: grp[0]) if single else (lambda grp: grp)
            ring_in, ring_out = [host_in, in2], [host_out, out2]
            t_start = eng.t
            st.step_ensemble(h, [wrap(ring_in[i & 1]) for i in range(2)], [wrap(ring_out[i & 1]) for i in range(2)])
            eng.t = t_start                   # members share the start time: rewind the clock for the next call
            barrier()
            t0 = time.perf_counter()
            Km = max(3, min(K, 20))           # members = timed steps (the pipeline fills once per call)
            st.step_ensemble(h, [wrap(ring_in[i & 1]) for i in range(Km)], [wrap(ring_out[i & 1]) for i in range(Km)])
            ens_s = time.perf_counter() - t0
            eng.t = t_start
            ok = True                         # against the serial call on the same inputs at the same time, bit for bit
            chk = [np.empty(s_, dtype=np.float64) for s_ in sizes]
            for b in range(2):
                one_step(ring_in[b], chk)
                eng.t = t_start
                ok = ok and all(np.array_equal(x, y) for x, y in zip(chk, ring_out[b]))
            e2e = {'value': n_total * Km / ens_s, 'unit': UNIT, 'h2d_bytes_per_step': 8 * sum(sizes),
                   'd2h_bytes_per_step': 8 * sum(sizes), 'steps': Km,
                   'note': 'public step_ensemble call over an ensemble of independent pinned-host states: per member the '
                           'state goes up, one step runs, the new state comes back; uploads, steps and read-backs of '
                           'consecutive members overlap on three streams; wall clock around the call (it synchronises)',
                   'check': 'bit-identical to the serial call' if ok else 'MISMATCH against the serial call',
                   'serial': {'value': serial['value'], 'note': serial['note']}}
            if not ok:
                e2e = dict(serial, ensemble_error='results differ from the serial call')
            in2 = out2 = ring_in = ring_out = chk = None
        except Exception as exc:               # keep the serial figure rather than lose the line
            e2e['ensemble_error'] = f'{type(exc).__name__}: {exc}'[:200]

    # ---- the same lattice with random edge weights: kernel bytes == algorithmic bytes (N=1 only) --------------------
    weighted = None
    if world == 1 and not args.no_weighted and not args.weighted and args.workload == 'heat2d':
        eng = st = fields = one_step = None
        prob.clear()
        pw = builder(product_api(), size, native=True, weighted=True)
        ew = pw['stepper']._ensure_engine()
        ew.run(ew.plan_steps(W, h))
        msw, _, _ = time_steps(ew, ctx, K, h, barrier, local, world, dist, torch, est_ms=est * 1.5)
        rw = roofline_of(pw, ew, K, h, msw / K)
        weighted = {'workload': pw['name'], 'value': pw['n_state'] * K / (msw * 1e-3), 'ms_per_step': msw / K,
                    'achieved': rw['achieved'], 'frac': rw['frac'], 'launch_ms': rw['launch_ms'],
                    'kernel_bytes_per_launch': rw['kernel_bytes_per_launch'], 'frac_kernel_bytes': rw['frac_kernel_bytes']}
        prob = pw
    roof['weighted_variant'] = weighted

    others = None
    if world == 1 and not args.no_workloads:
        eng = st = fields = one_step = host_in = host_out = None
        prob.clear()
        others = other_workloads(args, ctx, barrier, local, dist, torch, skip=args.workload)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.workload, args.cpu_size or cpu_size)

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong' if args.total else 'weak',
            'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD_NAME[0], 'scheme': SCHEME[0],
                       'l2': 'inputs exceed L2 (state + operators >> 126 MB)',
                       'parallelism': (f'{world} graph partitions (contiguous node ranges), halo exchange of boundary values '
                                       f'after every RK stage: ' + HALO[0]) if world > 1 else 'single GPU',
                       'dof_stage_evals_per_sec': 4 * value, 'build_seconds': t_build},
            'roofline': roof, 'cpu_baseline': cb, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
        }
        if multi is not None:
            out['multi_gpu'] = multi
        if others is not None:
            out['workloads'] = others
        emit(json.dumps(out))
    if world > 1:
        if getattr(eng, 'p2p', False):
            eng.sync()          # raises if an exchange of the e2e steps timed out
        dist.destroy_process_group()
        if multi is not None and not multi['ok']:
            raise SystemExit('bench.py: partitioned run is NOT bit-identical (see multi_gpu in the JSON line)')


WORKLOAD_NAME = [None]
BLOCKS = [None]          # --blocks pa,pb,pc: block decomposition of the 3-D grid instead of slabs
TOTAL = [False]          # --total: --size is the size of the WHOLE problem (strong scaling) instead of the size per rank
NO_PREROLL = [False]     # profiling runs (ncu replays every launch ~40 times): skip the 0.3 s of load for the clock sampler
SCHEME = ['rk4']
HALO = ['']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='heat2d', choices=list(WORKLOADS))
    ap.add_argument('--size', type=int, default=0)
    ap.add_argument('--cpu-size', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-weighted', action='store_true')
    ap.add_argument('--no-workloads', action='store_true', help='N = 1: skip the other BASELINE configurations after the headline')
    ap.add_argument('--no-pressure', action='store_true', help='N = 1: skip the Navier-Stokes workload with the per-step pressure solve')
    ap.add_argument('--no-slab-check', action='store_true', help='N > 1: skip the small partitioned-vs-one-GPU comparison')
    ap.add_argument('--weighted', action='store_true', help='heat2d: random edge weights on the main problem (kernel A/B runs)')
    ap.add_argument('--total', action='store_true', help='wave3d: --size is the whole grid, cut into --gpus slabs (strong scaling)')
    ap.add_argument('--blocks', default='', help='wave3d at N > 1: pa,pb,pc blocks along the label axes instead of slabs')
    ap.add_argument('--no-preroll', action='store_true', help='profiling runs under ncu: no untimed pre-roll before the timed steps')
    args = ap.parse_args()
    NO_PREROLL[0] = args.no_preroll
    TOTAL[0] = args.total
    BLOCKS[0] = tuple(int(x) for x in args.blocks.split(',')) if args.blocks else None
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
