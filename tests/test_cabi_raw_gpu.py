"""The C-ABI on its own: a heat problem assembled, stepped and read back through raw ctypes calls only -- no tracer, no
engine, no field classes -- the way a binding written inside the reference would drive libgds_b200.so.
Checked against the CPU oracle."""
import ctypes as C
import math

import numpy as np
import pytest

import cases
from adaptors import oracle_api, rel_l2

pytestmark = pytest.mark.gpu


def test_heat_rk4_through_the_raw_c_abi():
    from gds_b200 import _lib as L
    lib, ok = L.lib, L.check
    n, h, steps = 17, 5e-3, 25
    G = cases.g_grid(n, n)
    nodes = {v: i for i, v in enumerate(G.nodes())}                                  # gds.py:21
    eu = np.array([nodes[u] for u, v in G.edges()], dtype=np.int32)                  # gds.py:23
    ev = np.array([nodes[v] for u, v in G.edges()], dtype=np.int32)
    V, E = len(nodes), eu.size
    y0 = np.array([math.exp(-((x[0] - 8) ** 2 + (x[1] - 8) ** 2) / 9) for x in nodes])
    d_idx = np.array([i for x, i in nodes.items() if x[0] == 0 or x[0] == n - 1], dtype=np.int64)
    d_val = np.array([0.25 if x[0] == 0 else 0.0 for x in nodes if x[0] == 0 or x[0] == n - 1])
    y0[d_idx] = d_val

    ctx, graph, sysm, mask = L.c_vp(), L.c_vp(), L.c_vp(), L.c_vp()
    ok(lib.gds_ctx_create(0, C.byref(ctx)))
    fp = np.zeros(1, dtype=np.int64)
    ok(lib.gds_graph_create(ctx, V, E, L.ptr(eu), L.ptr(ev), None, 0, L.ptr(fp), None, C.byref(graph)))
    ok(lib.gds_system_create(ctx, graph, V, L.SCHEME_RK4, C.byref(sysm)))
    ok(lib.gds_mask_create(ctx, V, L.ptr(d_idx), d_idx.size, C.byref(mask)))
    # row program of  dy/dt = Z_D(L0 y):  N_LAP(stage vector), MASK_ZERO(Dirichlet rows); the RHS epilogue fuses the RK4 combine
    code = (C.c_uint32 * 2)(L.OP['N_LAP'] | (0 << 8), L.OP['MASK_ZERO'] | (0 << 8))
    vec = (L.VecRef * 1)()
    vec[0].kind, vec[0].buf, vec[0].offset = L.VEC_STAGE, None, 0
    masks = (L.c_vp * 1)(mask)
    d = L.PassDesc()
    d.domain, d.when = L.NODES, L.WHEN_STAGE
    d.n_code, d.code = 2, code
    d.n_vec, d.vec = 1, vec
    d.n_mask, d.mask = 1, masks
    d.epilogue, d.state_offset, d.dirichlet = L.EPI_RHS, 0, mask
    kind = C.c_int32()
    ok(lib.gds_pass_kernel_kind(C.byref(d), C.byref(kind)))
    assert kind.value == 1                                                            # the Laplacian kernels
    ok(lib.gds_system_add_pass(sysm, C.byref(d)))
    ok(lib.gds_system_set_dirichlet(sysm, L.ptr(d_idx), L.ptr(d_val), d_idx.size, 0))
    ok(lib.gds_system_finalize(sysm))
    ok(lib.gds_system_set_state(sysm, L.ptr(y0), 0, V))
    t = h * np.arange(steps, dtype=np.float64)
    hs = np.full(steps, h)
    ok(lib.gds_system_step(sysm, steps, L.ptr(t), L.ptr(hs), None))
    got = np.empty(V)
    ok(lib.gds_system_get_state(sysm, L.ptr(got), 0, V))
    launches, replays = L.c_i64(), L.c_i64()
    ok(lib.gds_ctx_counters(ctx, C.byref(launches), C.byref(replays)))
    kps, dev_bytes = L.c_i64(), L.c_i64()
    ok(lib.gds_system_info(sysm, C.byref(kps), C.byref(dev_bytes)))
    # 4 stages + the per-step tail kernel; a 17-wide lattice (odd row stride) runs its stages in the lane-per-row kernel,
    # whose last < 128 rows take a second launch
    assert kps.value in (5, 9) and dev_bytes.value > 0
    assert replays.value == steps and launches.value == steps * kps.value
    for hnd, fn in ((sysm, lib.gds_system_destroy), (mask, lib.gds_mask_destroy), (graph, lib.gds_graph_destroy),
                    (ctx, lib.gds_ctx_destroy)):
        ok(fn(hnd))

    api = oracle_api()
    T = api.node_gds(cases.g_grid(n, n))
    api.evolve(T, dydt=lambda tt, y: T.laplacian(y), max_step=h)
    T.set_initial(y0=lambda x: math.exp(-((x[0] - 8) ** 2 + (x[1] - 8) ** 2) / 9))
    T.set_constraints(dirichlet=lambda x: 0.25 if x[0] == 0 else (0.0 if x[0] == n - 1 else None))
    for _ in range(steps):
        T.step(h)
    assert rel_l2(got, T.y) <= 1e-10
