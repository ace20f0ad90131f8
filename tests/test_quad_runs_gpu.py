"""The quad Laplacian kernel gathers a run of consecutive neighbours (a lattice direction) as one access and moves its
streams as 256-bit accesses (kernels.cu, lapq_kernel<.., RUNS>).  It must produce the SAME BITS as the scalar-gather form
(GDS_B200_LAPQ_RUNS=0) on graphs that hit every branch -- aligned runs, +-1 runs through the row's own registers, even and
odd offsets, ragged last quads, graphs with no runs, weighted graphs -- and match the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from adaptors import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(tmp_path, tag, which, **env):
    out = str(tmp_path / f'{tag}.npz')
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, os.path.join(HERE, 'quad_runs_worker.py'), which, out], check=True, env=e, timeout=600)
    return np.load(out)


@pytest.mark.parametrize('force', ['', 'quad'])
def test_run_gathers_are_bit_identical_to_scalar_gathers(tmp_path, force):
    env = {'GDS_B200_LAP_KERNEL': force} if force else {}
    runs = _worker(tmp_path, 'runs' + force, 'product', GDS_B200_LAPQ_RUNS='1', **env)
    scalar = _worker(tmp_path, 'scalar' + force, 'product', GDS_B200_LAPQ_RUNS='0', **env)
    want = _worker(tmp_path, 'oracle' + force, 'oracle')
    for k in want.files:
        assert np.array_equal(runs[k], scalar[k]), k
        assert rel_l2(runs[k], want[k]) <= 1e-12, (k, rel_l2(runs[k], want[k]))
