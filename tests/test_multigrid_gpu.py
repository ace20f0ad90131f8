"""The multigrid-preconditioned solves on the device (gds_mg_*, gds_pcg_solve): the device V-cycle against the numpy model of
gds_b200/multigrid.py, and the Leray projection / the `lhs=` pressure solve with the preconditioner against the plain
conjugate-gradient path, the oracle and the reference's golden vectors."""
import numpy as np
import pytest

import cases
import gds_b200
from adaptors import oracle_api, product_api, rel_l2
from gds_b200 import multigrid as mg
from gds_b200.device import Buffer
from gds_b200.graph import LoweredGraph

pytestmark = pytest.mark.gpu

GRAPHS = {
    'hex': lambda: cases.g_hex(40, 40),
    'tri_weighted': lambda: cases.add_weights(cases.g_tri(50, 60), 3),
    'rgg': lambda: cases.g_rgg(3000, 0.04, 5),              # internally renumbered, variable degree
}


@pytest.mark.parametrize('gname', list(GRAPHS))
@pytest.mark.parametrize('sweeps', [(1, 1), (2, 2), (1, 2), (3, 1)])
def test_device_vcycle_equals_the_numpy_model(gname, sweeps):
    low = LoweredGraph.of(GRAPHS[gname]())
    dirichlet = np.arange(0, low.V, 41) if gname == 'tri_weighted' else ()
    levels = mg.build_levels(mg.masked_laplacian(low, dirichlet), coarse_size=120)
    assert len(levels) >= 3
    M = mg.Multigrid(low.ctx, levels, sweeps)
    info = M.info()
    assert info['levels'] == len(levels) and info['launches_per_cycle'] == (len(levels) - 1) * (sum(sweeps) + 3) + 1
    rng = np.random.default_rng(2)
    for _ in range(2):
        r = rng.standard_normal(low.V)
        rb, zb = Buffer(low.ctx, low.V, r), Buffer(low.ctx, low.V)
        M.apply(rb, zb)
        want = mg.vcycle_host(levels, r, pre=sweeps[0], post=sweeps[1])
        assert rel_l2(zb.download(), want) <= 1e-12
        assert np.array_equal(rb.download(), r)                      # the right-hand side is left alone


def test_single_level_hierarchy_is_a_direct_solve():
    low = LoweredGraph.of(cases.g_grid(9, 8))
    levels = mg.build_levels(mg.masked_laplacian(low))
    assert len(levels) == 1
    M = mg.Multigrid(low.ctx, levels)
    b = np.random.default_rng(5).standard_normal(low.V); b -= b.mean()
    rb, zb = Buffer(low.ctx, low.V, b), Buffer(low.ctx, low.V)
    M.apply(rb, zb)
    assert rel_l2(levels[0]['A'] @ zb.download(), b) <= 1e-10


@pytest.mark.parametrize('gname', list(GRAPHS))
def test_leray_projection_with_multigrid_matches_plain_cg(gname, monkeypatch):
    make = GRAPHS[gname]
    u = gds_b200.edge_gds(make())
    y = np.random.default_rng(8).standard_normal(u.ndim)
    monkeypatch.setenv('GDS_B200_MG', '0')
    z0, i0 = u.leray_project(y, return_info=True)
    i0 = dict(i0)
    monkeypatch.setenv('GDS_B200_MG', '1')
    z1, i1 = u.leray_project(y, return_info=True)
    assert i0['preconditioner'] is None and i1['preconditioner'] == 'multigrid'
    assert i1['relative_residual'] <= 1e-12 and i0['relative_residual'] <= 1e-12
    assert rel_l2(z1, z0) <= 1e-10
    assert i1['iterations'] <= 48 and i1['iterations'] * 2 <= i0['iterations'], (i1, i0)
    assert np.max(np.abs(u.div(z1))) <= 1e-10 * np.max(np.abs(u.div(y)))
    # the hierarchy is cached on the graph: the second projection reuses it and gives the same bits
    z2 = u.leray_project(y)
    assert np.array_equal(z1, z2) and len(u._low.__dict__['_mg_cache']) == 1


def test_leray_projection_with_multigrid_matches_the_dense_oracle(monkeypatch):
    monkeypatch.setenv('GDS_B200_MG', '1')
    for make in (lambda: cases.add_weights(cases.g_tri(9, 12), 4), lambda: cases.g_hex(7, 9), lambda: cases.g_rgg(300, 0.12, 1)):
        u, ref = gds_b200.edge_gds(make()), oracle_api().edge_gds(make())
        y = np.random.default_rng(8).standard_normal(u.ndim)
        z, info = u.leray_project(y, return_info=True)
        assert info['preconditioner'] == 'multigrid' and info['relative_residual'] <= 1e-12
        assert rel_l2(z, ref.leray_project(y)) <= 1e-10


def test_pressure_solve_with_multigrid_matches_plain_cg_and_the_oracle(monkeypatch):
    """`lhs=` pressure field (examples/fluid.py:14-38) on a lattice large enough for a three-level hierarchy"""
    kw = dict(m=24, n=40, nsteps=4, every=2)
    monkeypatch.setenv('GDS_B200_MG', '0')
    a = cases.case_ns_pressure(product_api(), **kw)
    monkeypatch.setenv('GDS_B200_MG', '1')
    monkeypatch.setattr(mg, 'COARSE_SIZE', 60)
    b = cases.case_ns_pressure(product_api(), **kw)
    for i in range(a['y_velocity'].shape[0]):
        assert rel_l2(b['y_velocity'][i], a['y_velocity'][i]) <= 1e-10
        assert rel_l2(b['y_pressure'][i], a['y_pressure'][i]) <= 1e-9
    # the small case of test_more_parity_gpu.py against the oracle's dense least squares, preconditioned
    c, d = cases.case_ns_pressure(product_api()), cases.case_ns_pressure(oracle_api())
    for i in range(d['y_velocity'].shape[0]):
        assert rel_l2(c['y_velocity'][i], d['y_velocity'][i]) <= 1e-9
        assert rel_l2(c['y_pressure'][i], d['y_pressure'][i]) <= 1e-8


def test_incomplete_or_mismatched_hierarchies_are_refused():
    low = LoweredGraph.of(cases.g_hex(20, 20))
    levels = mg.build_levels(mg.masked_laplacian(low), coarse_size=100)
    M = mg.Multigrid(low.ctx, levels)
    short = Buffer(low.ctx, low.V - 1)
    with pytest.raises(gds_b200.GdsError):
        M.apply(short, Buffer(low.ctx, low.V))
    same = Buffer(low.ctx, low.V)
    with pytest.raises(gds_b200.GdsError):
        M.apply(same, same)
    M2 = mg.Multigrid(low.ctx, levels[:-1])         # no coarsest level: refused at the first use
    with pytest.raises(gds_b200.GdsError):
        M2.apply(Buffer(low.ctx, low.V), Buffer(low.ctx, low.V))
