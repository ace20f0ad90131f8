#!/usr/bin/env python
"""bench.py -- DOF-updates/s of the gds stepping hot path on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat2d|hex_advdiff|ns_cavity|cml|wave3d] [--size S]
    python bench.py --impl reference ...     # the reference algorithm's CPU path (oracle port) on the host cores

A "step" is one fixed RK4 step (4 fused stage passes) -- or one map application for `cml` -- of the whole field.
Default workload: the heat law of BASELINE config 1 (`dydt = T.laplacian(y)`, static + time-varying Dirichlet) on
the 10^8-edge lattice `grid_2d_graph(7072, 7072)` the north-star roofline target is stated on.
`value` times K steps with the state resident in HBM (CUDA events on the library's stream, max over ranks);
`e2e` times the same steps through the public API with HOST state buffers: every step uploads the state from pinned
host memory, advances one step and reads the result back.  Inputs (>= 400 MB state + operators) exceed the 126 MB L2.
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
def heat2d_problem(api, n, native, world=1, rank=0, weighted=False):
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
        if world > 1:
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
    k_op = (16 + (32 if weighted else 0)) * V
    return dict(stepper=T, fields=[T], h=h, n_state=n_own, V=V, E=E, F=0,
                bytes_per_step=4 * m_op + 136 * V, bytes_per_launch=(4 * m_op + 136 * V) / 4.0,
                kernel_bytes_per_launch=(4 * k_op + 136 * V) / 4.0,
                dominant_class=0, kernel='lap4_kernel<unit weights>' if not weighted else 'lap4_kernel<weighted>',
                name=f'heat RK4 h=1e-2 on grid_2d_graph({m},{n}) (V={m * n}, E={2 * m * n - m - n}), Dirichlet static+time-varying'
                     + (', random edge weights' if weighted else '')
                     + (f', {world} slabs of {n}x{n} with per-stage halo exchange' if world > 1 else ''))


WORKLOADS = {'heat2d': (heat2d_problem, 7072, 1024)}   # name -> (builder, default GPU size per rank, CPU sample size)


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
def cpu_run(workload, size, steps, warmup):
    from adaptors import oracle_api
    builder = WORKLOADS[workload][0]
    prob = builder(oracle_api(), size, native=False)
    st, h = prob['stepper'], prob['h']
    for _ in range(warmup):
        st.step(h)
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(h)
    el = time.perf_counter() - t0
    return prob, el


def cpu_baseline(workload, size, budget_s=12.0):
    """bounded sample of the same law on the host: oracle port (scipy CSR SpMV, single thread)"""
    prob, el1 = cpu_run(workload, size, steps=3, warmup=1)
    steps = int(max(3, min(400, budget_s / max(el1 / 3, 1e-6))))
    st, h = prob['stepper'], prob['h']
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(h)
    el = time.perf_counter() - t0
    return {'value': prob['n_state'] * steps / el, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': f'{prob["name"]}; {steps} steps in {el:.2f} s (oracle/gds_oracle.py, scipy CSR, 1 thread of '
                      f'{os.cpu_count()} host cores)'}


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    workload = args.workload
    size = args.cpu_size or WORKLOADS[workload][2]
    prob, el = cpu_run(workload, size, args.steps, args.warmup)
    v = prob['n_state'] * args.steps / el
    cb = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
          'sample': f'{prob["name"]}; oracle/gds_oracle.py (numpy/scipy restatement of the reference, 1 thread of '
                    f'{os.cpu_count()} host cores: scipy CSR SpMV is single-threaded)'}
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': prob['name'], 'note': 'bounded CPU sample of the GPU arm\'s law'},
        'cpu_baseline': cb,
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_steps(eng, ctx, plan, barrier, device, world, dist, torch):
    """K steps with the state resident in HBM: CUDA events on the library's stream, max over ranks"""
    launches0, _ = ctx.counters()
    barrier()
    with ClockSampler(device) as clk:
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


def roofline_of(prob, eng, K, h, ms_step):
    """per-launch time of the dominant kernel: a CUDA-event pair around every launch, eager replay of the same steps"""
    cls_ms, cls_n = eng.run_timed(eng.plan_steps(min(K, 64), h))
    dom = prob['dominant_class']
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
    prob = builder(product_api(), size, native=True, world=world, rank=rank)
    WORKLOAD_NAME[0] = prob['name']
    st, h, K, W = prob['stepper'], prob['h'], args.steps, args.warmup
    eng = st._ensure_engine()
    ctx = eng.ctx
    t_build = time.perf_counter() - t_build

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- kernel-resident throughput: K steps, state in HBM ----------------------------------------------------
    eng.run(eng.plan_steps(W, h))
    ms, launches, clocks = time_steps(eng, ctx, eng.plan_steps(K, h), barrier, local, world, dist, torch)
    n_total = prob['n_state'] * world
    value = n_total * K / (ms * 1e-3)
    roof = roofline_of(prob, eng, K, h, ms / K)
    try:   # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            roof['traffic'] = json.load(fh).get(args.workload, {}).get(str(size))
    except Exception:
        pass

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    n = prob['n_state']
    host_in = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    host_out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    f = prob['fields'][0]
    eng.read_into(f, host_in)
    Ke = max(3, min(K, 10))
    for _ in range(2):
        eng.write(f, host_in); st.step(h); eng.read_into(f, host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        eng.write(f, host_in)          # H2D of the step's input state (pinned)
        st.step(h)                     # public API call
        eng.read_into(f, host_out)     # D2H of the step's result (synchronises)
        host_in, host_out = host_out, host_in
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {'value': n_total * Ke / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': 8 * n, 'd2h_bytes_per_step': 8 * n,
           'steps': Ke, 'note': 'per step: pinned-host state upload, one public step() call, state read-back; wall clock'}

    # ---- the same lattice with random edge weights: kernel bytes == algorithmic bytes (N=1 only) --------------------
    weighted = None
    if world == 1 and not args.no_weighted:
        del eng, st, f
        prob.clear()
        pw = builder(product_api(), size, native=True, weighted=True)
        ew = pw['stepper']._ensure_engine()
        ew.run(ew.plan_steps(W, h))
        msw, _, _ = time_steps(ew, ctx, ew.plan_steps(K, h), barrier, local, world, dist, torch)
        rw = roofline_of(pw, ew, K, h, msw / K)
        weighted = {'workload': pw['name'], 'value': pw['n_state'] * K / (msw * 1e-3), 'ms_per_step': msw / K,
                    'achieved': rw['achieved'], 'frac': rw['frac'], 'launch_ms': rw['launch_ms'],
                    'kernel_bytes_per_launch': rw['kernel_bytes_per_launch'], 'frac_kernel_bytes': rw['frac_kernel_bytes']}
        prob = pw
    roof['weighted_variant'] = weighted

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.workload, args.cpu_size or cpu_size)

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': roof.pop('workload_name', None) or WORKLOAD_NAME[0], 'scheme': 'rk4',
                       'l2': 'inputs exceed L2 (state + operators >> 126 MB)',
                       'parallelism': (f'{world} graph partitions (contiguous slabs), halo exchange of boundary node '
                                       f'values after every RK stage over NCCL send/recv') if world > 1 else 'single GPU',
                       'dof_stage_evals_per_sec': 4 * value, 'build_seconds': t_build},
            'roofline': roof, 'cpu_baseline': cb, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


WORKLOAD_NAME = [None]


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
