"""The committed golden vectors ARE the outputs of the unmodified reference: where /root/reference is mounted (the build
container; not the GPU box) a sample of them is regenerated in-process through tests/golden/make_golden.py's adaptor and
compared with the committed files.  Skipped when the reference is absent."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/gds'), reason='reference not mounted')


@pytest.fixture(scope='module')
def ref_api():
    argv, sys.argv = sys.argv, ['make_golden']
    try:
        spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLD, 'make_golden.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod.RefAPI


def _same(out, path):
    gold = np.load(path)
    assert set(out) == set(gold.files)
    for k in gold.files:
        a, b = np.asarray(out[k]), gold[k]
        assert a.shape == b.shape, k
        if a.dtype.kind in 'fc':
            assert np.allclose(a, b, rtol=1e-13, atol=1e-15), k     # same code, same machine class: last-bit BLAS noise only
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize('gname', ['grid_6x7', 'hex_3x4'])
def test_operator_golden_is_fresh(ref_api, gname):
    _same(cases.operator_outputs(ref_api, gname), os.path.join(GOLD, f'ops_{gname}.npz'))


@pytest.mark.parametrize('cname', ['order2_dirichlet', 'map_time_varying_dirichlet', 'face_diffusion'])
def test_quirk_golden_is_fresh(ref_api, cname):
    _same(cases.QUIRK_CASES[cname](ref_api), os.path.join(GOLD, f'quirk_{cname}.npz'))


def test_trajectory_golden_is_fresh(ref_api):
    _same(cases.TRAJ_CASES['heat_c1_euler_small'](ref_api), os.path.join(GOLD, 'traj_heat_c1_euler_small.npz'))


def test_jacobian_and_ragged_golden_are_fresh(ref_api):
    _same(cases.jacobian_outputs(ref_api, 'hex_3x4'), os.path.join(GOLD, 'jac_hex_3x4.npz'))
    _same(cases.run_ops(ref_api, cases.RAGGED['star_70_weighted']()), os.path.join(GOLD, 'ragged_star_70_weighted.npz'))
