"""bench.py's CPU arm keeps the driver's contract: ONE JSON line on stdout with the metric of BASELINE.json, the
cpu_baseline object describing the run and an e2e object equal to the line's own value (no GPU involved)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=600,
                       env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]                      # library chatter goes to stderr
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = run_bench('--impl', 'reference', '--steps', '2', '--warmup', '1', '--no-cpu')
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['steps'] == 2 and d['warmup'] == 1
    assert d['metric'] == 'dof_updates_per_sec' and d['unit'] == 'DOF-updates/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None and d['dtype'] == 'f64'
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] == 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'grid_2d_graph' in d['config']['workload']


def test_other_ranks_of_the_reference_arm_exit_without_work():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '2'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT,
                       env=dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1'))
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '2'], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
