"""Parity of the CUDA product (gds_b200, through the C-ABI) against
  * the golden vectors produced by the UNMODIFIED reference (tests/golden/*.npz), and
  * the CPU oracle on the same seeded inputs.
Index maps bit-exact; float64 results within 1e-10 relative L2 (north_star tolerance)."""
import os

import numpy as np
import pytest

import cases
from adaptors import oracle_api, product_api, rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 1e-10
INT_KEYS = ('map_edges', 'map_faces_ptr', 'map_faces_nodes', 'n_faces', 'map_nodes_repr')


@pytest.mark.parametrize('gname', list(cases.OP_GRAPHS))
def test_operators_match_reference(gname):
    gold = np.load(os.path.join(GOLD, f'ops_{gname}.npz'))
    out = cases.operator_outputs(product_api(), gname)
    assert set(out) == set(gold.files)
    for k in gold.files:
        if k in INT_KEYS:
            assert np.array_equal(out[k], gold[k]), k
        elif k.startswith('in_') or k == 'edge_weights':
            assert np.array_equal(out[k], gold[k]), k
        else:
            assert rel_l2(out[k], gold[k]) <= TOL, (k, rel_l2(out[k], gold[k]))


@pytest.mark.parametrize('gname', list(cases.JAC_GRAPHS))
def test_advection_jacobian(gname):
    """edge_gds.advect_jac / advect_jvp (gds.py:391-395).  The reference differentiates through float32 copies of its
    operators (gds.py:270-271), so its Jacobian carries ~1e-8 of rounding wherever a weight or 1/(deg-1) is not a dyadic
    number; the device path computes in float64.  Bars: 1e-10 against the full-precision oracle on every graph, 1e-10
    against the reference's golden Jacobian on the graph whose operator entries are exact in float32 (unit-weight
    hexagons), 1e-6 (float32 resolution) on the others."""
    gold = np.load(os.path.join(GOLD, f'jac_{gname}.npz'))
    out = cases.jacobian_outputs(product_api(), gname)
    assert np.array_equal(out['in_y'], gold['in_y'])
    y = gold['in_y']
    ed = oracle_api().edge_gds(cases.OP_GRAPHS[gname]())
    assert rel_l2(out['jac'], ed.advect_jac(y, single_precision_operators=False)) <= TOL
    assert rel_l2(out['jac'], gold['jac']) <= (TOL if gname == 'hex_3x4' else 1e-6)
    pd = product_api().edge_gds(cases.OP_GRAPHS[gname]())
    v = np.random.default_rng(6).standard_normal(y.size)
    assert rel_l2(pd.advect_jvp(v, y), ed.advect_jvp(v, y)) <= TOL


@pytest.mark.parametrize('cname', list(cases.TRAJ_CASES))
def test_trajectories_match_reference(cname):
    gold = np.load(os.path.join(GOLD, f'traj_{cname}.npz'))
    out = cases.TRAJ_CASES[cname](product_api())
    assert set(out) == set(gold.files)
    assert np.allclose(out['t'], gold['t'], rtol=0, atol=1e-12)
    for k in gold.files:
        if k == 't':
            continue
        for i in range(gold[k].shape[0]):
            err = rel_l2(out[k][i], gold[k][i])
            assert err <= TOL, (cname, k, i, err)


@pytest.mark.parametrize('cname,kw', [
    ('hex_advdiff_c2', dict(m=20, n=24, nsteps=30, every=30)),
    ('ns_cavity_c3', dict(m=20, n=30, nsteps=30, every=30)),
    ('wave_c5', dict(dims=(12, 13, 14), nsteps=40, every=40)),
    ('cml_c4', dict(n=5000, iters=20, every=20)),
])
def test_trajectories_match_oracle_larger(cname, kw):
    """sizes the dense reference cannot hold: CUDA product vs the sparse CPU oracle"""
    fn = {'hex_advdiff_c2': cases.case_hex_advdiff_c2, 'ns_cavity_c3': cases.case_ns_cavity_c3,
          'wave_c5': cases.case_wave_c5, 'cml_c4': cases.case_cml_c4}[cname]
    want = fn(oracle_api(), **kw)
    got = fn(product_api(), **kw)
    for k in want:
        if k == 't':
            assert np.allclose(got[k], want[k], rtol=0, atol=1e-12)
            continue
        for i in range(want[k].shape[0]):
            err = rel_l2(got[k][i], want[k][i])
            assert err <= TOL, (cname, k, i, err)
