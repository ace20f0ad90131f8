"""The partitioned stepping path (stage-by-stage enqueue + halo pack / exchange / unpack) on ONE GPU: two ranks share
cuda:0 and exchange through gloo (NCCL cannot put two ranks on one device; the transport is the only difference from
the multi-GPU run).  Assembled results must match the CPU oracle of the undistributed problem to 1e-10, and the
undistributed CUDA run bit for bit (same summation order in every owned row)."""
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
}


def _rd_grid(api):
    G = cases.g_grid(14, 10)
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: 0.3 * T.laplacian(y) + y - y ** 3, max_step=5e-3)
    y0 = np.random.default_rng(21).uniform(-1, 1, 140)
    T.set_initial(y0=lambda x: y0[x[0] * 10 + x[1]])
    return cases._snap(T, {'T': T}, 5e-3, 20, 10)


def _worker(rank, world, port, name, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), GDS_B200_DEVICE='0')
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import gds_b200
        fn, align, hops = CASES[name]
        cases.GRAPH_HOOK = lambda G: gds_b200.distribute(G, rank, world, align=align, hops=hops)
        out = fn(product_api())
        ret[rank] = {k: np.asarray(v) for k, v in out.items()}
    finally:
        cases.GRAPH_HOOK = None
        dist.destroy_process_group()


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_matches_undistributed(name, world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + hash(name) + world) % 1500
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, name, ret), nprocs=world, join=True)
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
