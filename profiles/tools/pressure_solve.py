#!/usr/bin/env python
"""Time the node-domain solve of the path (Leray projection: B1 B1^T phi = B1 y, gds/gds.py:249,414-419) on a BASELINE-size
lattice with and without the multigrid preconditioner.  One GPU.

    python profiles/tools/pressure_solve.py [hexagonal|triangular] [size] [--no-cg] > gpurun_out/<tag>_pressure_solve.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gds_b200  # noqa: E402
from gds_b200.graph import NativeGraph  # noqa: E402
from gds_b200.multigrid import Multigrid  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    kind = args[0] if args else 'hexagonal'
    size = int(args[1]) if len(args) > 1 else 2000
    G = NativeGraph(kind, size, size)
    u = gds_b200.edge_gds(G)
    low = u._low
    y = np.random.default_rng(0).standard_normal(u.ndim)
    out = {'graph': f'{kind}_lattice_graph({size},{size})', 'V': low.V, 'E': low.E, 'rtol': 1e-10}
    os.environ['GDS_B200_MG'] = '1'
    t0 = time.perf_counter()
    M = Multigrid.for_graph(low)
    low.ctx.sync()
    out['hierarchy_seconds'] = time.perf_counter() - t0
    out['hierarchy'] = M.info()
    for rep in range(2):                     # the second call is the steady state (buffers and passes exist)
        t0 = time.perf_counter()
        z, info = u.leray_project(y, rtol=1e-10, return_info=True)
        out['multigrid'] = dict(info, seconds=time.perf_counter() - t0)
    div0 = float(np.max(np.abs(u.div(y))))
    out['multigrid']['max_div_after_over_before'] = float(np.max(np.abs(u.div(z)))) / div0
    if '--no-cg' not in sys.argv:
        os.environ['GDS_B200_MG'] = '0'
        t0 = time.perf_counter()
        z0, info0 = u.leray_project(y, rtol=1e-10, max_iter=40000, return_info=True)
        out['plain_cg'] = dict(info0, seconds=time.perf_counter() - t0)
        out['rel_l2_multigrid_vs_plain'] = float(np.linalg.norm(z - z0) / np.linalg.norm(z0))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
