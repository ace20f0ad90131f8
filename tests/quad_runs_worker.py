"""Helper of test_quad_runs_gpu.py (not collected): steps a set of node-Laplacian problems on the device and saves the
states.  Run once per kernel setting (GDS_B200_LAPQ_RUNS / GDS_B200_LAP_KERNEL are read when the library first launches)."""
import os
import sys

import networkx as nx
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def _plain(G):
    G.faces = []
    return G


def graphs():
    import cases
    yield 'grid_64x64', cases.g_grid(64, 64)          # +-64: 32-byte aligned runs; +-1: register runs
    yield 'grid_37x41', cases.g_grid(37, 41)          # +-41: odd offset (scalar run); ragged last quad (1517 rows)
    yield 'grid_30x42', cases.g_grid(30, 42)          # +-42: even, not a multiple of 4 (two 128-bit loads)
    yield 'grid3d_6x10x12', cases.g_grid3((6, 10, 12))
    yield 'path_1000', _plain(nx.path_graph(1000))
    yield 'cycle_515', _plain(nx.cycle_graph(515))
    yield 'hex_9x11', cases.g_hex(9, 11)
    yield 'rgg_800', cases.g_rgg(800, 0.08, 5)        # no runs at all
    yield 'grid_64x64_w', cases.add_weights(cases.g_grid(64, 64), 21)


def run(api, name, G):
    T = api.node_gds(G)
    api.evolve(T, dydt=lambda t, y: 0.3 * T.laplacian(y), max_step=1e-2)
    y0 = np.random.default_rng(7).standard_normal(T.ndim)
    T.set_initial(y0=lambda x: y0[T.X[x]])
    nodes = list(T.X.keys())
    T.set_constraints(dirichlet={nodes[i]: 0.5 for i in range(0, len(nodes), 97)})
    out = [np.array(T.laplacian(y0), copy=True)]
    for _ in range(3):
        T.step(0.02)
        out.append(np.array(T.y, copy=True))
    return np.array(out)


def main(which, path):
    from adaptors import oracle_api, product_api
    api = product_api() if which == 'product' else oracle_api()
    np.savez(path, **{name: run(api, name, G) for name, G in graphs()})


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
