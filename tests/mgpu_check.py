"""Multi-GPU check, run under torchrun on N >= 2 real GPUs (not collected by pytest: one GPU per rank is needed):

    torchrun --nproc-per-node N tests/mgpu_check.py

Every case runs partitioned with the peer-to-peer halo (default on NCCL) and with the host-driven NCCL send/recv halo;
rank 0 also runs the undistributed problem on its GPU; the assembled results must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    os.environ['GDS_B200_DEVICE'] = str(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import cases
    import gds_b200
    from adaptors import product_api
    CASES = {
        'heat': (lambda api: cases.case_heat_c1(api, n=96, nsteps=40, every=20), 96, 1),
        # slabs of > 32767 owned rows on up to 8 ranks: the ghost lattice rows sit behind the 16-bit index escape, and the
        # peer-to-peer transport splits every stage into boundary rows -> push || interior rows
        'heat_big': (lambda api: cases.case_heat_c1(api, n=528, nsteps=10, every=5), 528, 1),
        'wave_big': (lambda api: cases.case_wave_c5(api, dims=(16, 16, 128), nsteps=10, every=5), 256, 1),
        'heat_mixed': (cases.case_heat_mixed, 9, 1),
        'wave': (lambda api: cases.case_wave_c5(api, dims=(7, 8, 24), nsteps=20, every=10), 56, 1),
        'cml': (lambda api: cases.case_cml_c4(api, n=3000, iters=12, every=6), 1, 1),
        # edge / face fields: deep partitions (several ghost layers, owned rows not a prefix)
        'hex_advdiff': (lambda api: cases.case_hex_advdiff_c2(api, m=8, n=10, nsteps=12, every=6), 1, 4),
        'ns_cavity': (lambda api: cases.case_ns_cavity_c3(api, m=8, n=12, nsteps=12, every=6), 1, 2),
    }
    ok = True
    for name, (fn, align, hops) in CASES.items():
        results = {}
        for mode in ('p2p', 'nccl'):
            os.environ['GDS_B200_HALO'] = mode
            cases.GRAPH_HOOK = lambda G: gds_b200.distribute(G, rank, world, align=align, hops=hops)
            out = fn(product_api())
            cases.GRAPH_HOOK = None
            results[mode] = {k: np.asarray(v) for k, v in out.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        if rank == 0:
            single = fn(product_api())
            for mode in ('p2p', 'nccl'):
                for k in [k for k in single if k.startswith('y_')]:
                    f = k[2:]
                    got = np.full_like(single[k], np.nan)
                    for r in range(world):
                        own, gid = gathered[r][mode][f'own_{f}'], gathered[r][mode][f'gid_{f}']
                        got[:, gid[own]] = gathered[r][mode][k][:, own]
                    same = np.array_equal(got, single[k])
                    ok = ok and same
                    print(f'{name:12s} {mode:5s} {k}: bit-identical to one GPU: {same}'
                          + ('' if same else f' (max abs diff {np.nanmax(np.abs(got - single[k])):.3e})'), flush=True)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
