"""The partitioned stepping path on ONE GPU: two or three ranks (processes) share cuda:0.
  * transport 'host': stage-by-stage enqueue + halo pack / exchange / unpack, exchanged through gloo (NCCL cannot put two
    ranks on one device);
  * transport 'p2p': the production path of the multi-GPU run -- the push / wait kernels fused into the captured step
    store straight into the neighbour's state buffers, mapped through CUDA IPC (which works between processes that share a
    device; the GPU time-slices the ranks, so a spinning wait kernel is preempted for the rank it waits for).
Assembled results must match the CPU oracle of the undistributed problem to 1e-10, and the undistributed CUDA run bit for
bit (same summation order in every owned row)."""
import os

import numpy as np
import pytest

import cases
from adaptors import oracle_api, product_api, rel_l2

pytestmark = pytest.mark.gpu

# name -> (case, alignment of the cuts, ghost layers).  The last two carry edge fields and two-hop laws: they need the deep
# partition (hex faces: 6 nodes -> 4 layers; triangles: 2).
CASES = {
    'heat': (lambda api: cases.case_heat_c1(api, n=24, nsteps=30, every=15), 24, 1),
    'heat_mixed': (cases.case_heat_mixed, 9, 1),
    'reaction_diffusion_grid': (lambda api: _rd_grid(api), 10, 1),
    'wave': (lambda api: cases.case_wave_c5(api, dims=(5, 6, 9), nsteps=20, every=10), 30, 1),
    'cml': (lambda api: cases.case_cml_c4(api, n=600, iters=12, every=6), 1, 1),
    'hex_advdiff': (lambda api: cases.case_hex_advdiff_c2(api, m=8, n=10, nsteps=12, every=6), 1, 4),
    'ns_cavity': (lambda api: cases.case_ns_cavity_c3(api, m=8, n=12, nsteps=12, every=6), 1, 2),
    # slabs of >= 1024 owned rows: the p2p transport splits every stage into boundary rows -> push || interior rows
    'heat_overlap': (lambda api: cases.case_heat_c1(api, n=64, nsteps=12, every=6), 64, 1),
    'wave_overlap': (lambda api: cases.case_wave_c5(api, dims=(8, 8, 48), nsteps=10, every=5), 64, 1),
}


def _rd_grid(api):
    G = cases.g_grid(14, 10)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: 0.3 * T.laplacian(y) + y - y ** 3, max_step=5e-3)
    y0 = np.random.default_rng(21).uniform(-1, 1, 140)
    T.set_initial(y0=lambda x: y0[x[0] * 10 + x[1]])
    return cases._snap(T, {'T': T}, 5e-3, 20, 10)


def _worker(rank, world, port, name, transport, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), GDS_B200_DEVICE='0', GDS_B200_HALO=transport)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import gds_b200
        from gds_b200 import engine as E
        fn, align, hops = CASES[name]
        cases.GRAPH_HOOK = lambda G: gds_b200.distribute(G, rank, world, align=align, hops=hops)
        seen = []
        build = E.StepEngine._build

        def spy(self):
            build(self)
            seen.append((bool(self.p2p), self.overlap))
        E.StepEngine._build = spy
        out = fn(product_api())
        ret[rank] = {k: np.asarray(v) for k, v in out.items()}
        ret[f'engine{rank}'] = seen
    finally:
        cases.GRAPH_HOOK = None
        dist.destroy_process_group()


# host transport: every small case on 2 and 3 ranks; p2p: every case on 2 ranks, two of them on 3 (middle rank: 2 neighbours)
RUNS = ([(n, w, 'host') for n in CASES if not n.endswith('_overlap') for w in (2, 3)]
        + [(n, 2, 'p2p') for n in CASES] + [('heat', 3, 'p2p'), ('hex_advdiff', 3, 'p2p')])


@pytest.mark.parametrize('name,world,transport', RUNS)
def test_partitioned_matches_undistributed(name, world, transport):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + hash(name) + world + (7 if transport == 'p2p' else 0)) % 1500
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, name, transport, ret), nprocs=world, join=True)
    for r in range(world):
        # the transport asked for is the one that ran; the large slabs took the overlap split
        assert all(p2p == (transport == 'p2p') for p2p, _ in ret[f'engine{r}']), ret[f'engine{r}']
        if transport == 'p2p' and name.endswith('_overlap') and world == 2:
            assert all(ov is not None for _, ov in ret[f'engine{r}']), ret[f'engine{r}']
    fn = CASES[name][0]
    want = fn(oracle_api())
    single = fn(product_api())
    keys = [k for k in want if k.startswith('y_')]
    for k in keys:
        f = k[2:]
        got = np.full_like(want[k], np.nan)
        for r in range(world):
            own, gid = ret[r][f'own_{f}'], ret[r][f'gid_{f}']
            got[:, gid[own]] = ret[r][k][:, own]
            # ghost rows hold their owner's values after every step
            assert np.array_equal(ret[r][k][:, ~own], single[k][:, gid[~own]])
        assert np.all(np.isfinite(got))
        assert np.array_equal(got, single[k]), (name, k, 'partitioned result differs from the one-GPU result')
        for i in range(want[k].shape[0]):
            assert rel_l2(got[i], want[k][i]) <= 1e-10


def _timeout_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), GDS_B200_DEVICE='0', GDS_B200_HALO='p2p',
                      GDS_B200_P2P_TIMEOUT_MS='1500')
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import gds_b200
        G = gds_b200.distribute(cases.g_grid(24, 24), rank, world, align=24)
        T = gds_b200.node_gds(G)
        T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2, integrator=gds_b200.Integrators.rk4)
        T.set_initial(y0=lambda x: float(x[0] + x[1]))
        eng = T._ensure_engine()
        assert eng.p2p
        # rank 1 takes one step fewer: rank 0's last step waits for values that never come
        eng.run(eng.plan_steps(3 if rank == 0 else 2, 1e-2))
        try:
            eng.sync()
            ret[rank] = 'ok'
        except gds_b200.GdsError as e:
            ret[rank] = str(e)
        if rank == 0:          # once raised, the flag stays up: every later call refuses to go on
            try:
                eng.run(eng.plan_steps(1, 1e-2))
                ret['again'] = 'ok'
            except gds_b200.GdsError as e:
                ret['again'] = str(e)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_unmatched_exchange_raises_instead_of_returning_stale_ghosts():
    """a rank whose neighbour stops delivering must not hand out its state: the wait kernel gives up after the timeout
    (1.5 s here, 20 s by default), later exchanges return at once, and the next host call on the engine raises"""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 977) % 1500
    ret = mp.Manager().dict()
    mp.spawn(_timeout_worker, args=(2, port, ret), nprocs=2, join=True)
    assert 'timed out waiting for rank 1' in ret[0], ret[0]
    assert 'timed out' in ret['again'], ret['again']
    assert ret[1] == 'ok'


def _native_grid_problem(dist_args=None):
    """heat (stage-state Laplacian, time-varying Dirichlet plane) and wave (order 2, accepted-state Laplacian: collapsed)
    on a native 3-D grid; returns {name: (global ids of the owned rows, their values)}"""
    import gds_b200
    out = {}
    dims = (10, 12, 14)                       # grid_graph(dim=[10, 12, 14]): labels (a, b, c) with a < 14, b < 12, c < 10
    A, B, C = dims[2], dims[1], dims[0]
    for name in ('heat', 'wave'):
        G = gds_b200.NativeGraph('grid', *dims)
        if dist_args is not None:
            gds_b200.distribute(G, *dist_args[0], **dist_args[1])
        T = gds_b200.node_gds(G)
        gid = np.asarray(T.global_ids)
        a, b, c = gid // (B * C), (gid // C) % B, gid % C
        if name == 'heat':
            T.set_evolution(dydt=lambda t, y: 0.4 * T.laplacian(y), max_step=2e-2, integrator=gds_b200.Integrators.rk4)
            T.set_initial(y0=np.sin(0.3 * a) + 0.1 * b - 0.05 * c)
            plane = np.arange(B * C)             # a == 0
            T.set_constraints(dirichlet=gds_b200.BoundarySpec(plane, fun=lambda t: np.cos(t) * np.ones(plane.size)))
        else:
            T.set_evolution(dydt=lambda t, y: 0.5 * T.laplacian(), order=2, max_step=5e-2, integrator=gds_b200.Integrators.rk4)
            T.set_initial(y0=np.exp(-((a - 6.5) ** 2 + (b - 5.5) ** 2 + (c - 4.5) ** 2) / 8.0))
        for _ in range(4):
            T.step(0.1)
        own = np.asarray(T.owned_mask)
        out[name] = (gid[own], np.asarray(T.y)[own], T._ensure_engine().overlap, bool(T._ensure_engine().p2p))
    return out


def _block_worker(rank, world, port, blocks, transport, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), GDS_B200_DEVICE='0', GDS_B200_HALO=transport)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = _native_grid_problem(((rank, world), dict(blocks=blocks)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('blocks,transport', [((2, 2, 1), 'p2p'), ((2, 1, 2), 'host'), ((2, 2, 2), 'p2p')])
def test_block_decomposition_matches_undistributed(blocks, transport):
    """blocks of a 3-D grid (BlockPartition: up to 6 face neighbours, boundary rows first) against the one-GPU run, bit for
    bit; the peer-to-peer transport hides the exchange behind the interior rows here too"""
    import torch.multiprocessing as mp
    world = int(np.prod(blocks))
    port = 29500 + (os.getpid() + 31 * world + (5 if transport == 'p2p' else 0)) % 1500
    ret = mp.Manager().dict()
    mp.spawn(_block_worker, args=(world, port, blocks, transport, ret), nprocs=world, join=True)
    single = _native_grid_problem()
    for name, (gid1, y1, _, _) in single.items():
        got = np.full_like(y1, np.nan)
        for r in range(world):
            gid, y, overlap, p2p = ret[r][name]
            got[gid] = y
            assert p2p == (transport == 'p2p')
        assert np.array_equal(gid1, np.arange(y1.size)) and np.array_equal(got, y1), name
