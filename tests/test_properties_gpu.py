"""Edge cases (empty / tiny / ragged graphs, sizes off the 32- and 128-row boundaries) against the CPU oracle, and
size-independent properties at BASELINE.json's full sizes (where the oracle would take too long):
conservation, linearity, curl o grad = 0, 2-homogeneity of the edge self-advection."""
import math
import os

import networkx as nx
import numpy as np
import pytest

from adaptors import oracle_api, product_api, rel_l2

pytestmark = pytest.mark.gpu


import cases
from cases import RAGGED, run_ops


@pytest.mark.parametrize('name', list(RAGGED))
def test_ragged_and_tiny_graphs_match_oracle(name):
    want = run_ops(oracle_api(), RAGGED[name]())
    got = run_ops(product_api(), RAGGED[name]())
    assert set(want) == set(got)
    for k in want:
        assert np.shape(got[k]) == np.shape(want[k]), k
        assert rel_l2(got[k], want[k]) <= 1e-10, (name, k, rel_l2(got[k], want[k]))
    gold_file = os.path.join(os.path.dirname(__file__), 'golden', f'ragged_{name}.npz')
    if os.path.exists(gold_file):      # the unmodified reference on the same graph (it cannot construct all of them)
        gold = np.load(gold_file)
        assert set(gold.files) == set(got)
        for k in gold.files:
            assert rel_l2(got[k], gold[k]) <= 1e-10, (name, k, rel_l2(got[k], gold[k]))




@pytest.mark.parametrize('gname', list(cases.fuzz_graphs()))
def test_irregular_graphs_match_oracle(gname):
    """random / irregular graphs (the oracle equals the live reference on them, tests/test_golden_fresh_cpu.py)"""
    want = run_ops(oracle_api(), cases.fuzz_graphs()[gname])
    got = run_ops(product_api(), cases.fuzz_graphs()[gname])
    assert set(want) == set(got)
    for k in want:
        assert np.shape(got[k]) == np.shape(want[k]), k
        assert rel_l2(got[k], want[k]) <= 1e-10, (gname, k, rel_l2(got[k], want[k]))


def test_matrix_right_hand_side():
    """operators accept a matrix of columns, e.g. laplacian(np.eye(n)) to extract the operator (examples/fluid.py:350)"""
    G = cases.g_hex(3, 4)
    a, b = product_api().edge_gds(G), oracle_api().edge_gds(cases.g_hex(3, 4))
    I = np.eye(a.ndim)
    Lg, Lo = a.laplacian(I), np.stack([b.laplacian(I[:, j]) for j in range(a.ndim)], axis=1)
    assert Lg.shape == Lo.shape and np.max(np.abs(Lg - Lo)) <= 1e-12


# ------------------------------------------------------------------------------------------------
# full BASELINE sizes: properties instead of an oracle run
# ------------------------------------------------------------------------------------------------
def test_heat_1e8_edges_conservation_and_linearity():
    """C1 law on grid_2d_graph(7072, 7072) (10^8 edges): L0 has zero column sums, so RK4 on dy/dt = L0 y conserves
    sum(y) (no Dirichlet rows); and the stepping map is linear"""
    import gds_b200
    n = 7072
    G = gds_b200.NativeGraph('grid_2d', n, n)
    T = gds_b200.node_gds(G)
    T.set_evolution(dydt=lambda t, y: T.laplacian(y), max_step=1e-2, integrator=gds_b200.Integrators.rk4)
    rng = np.random.default_rng(0)
    x, z = rng.random(n * n), rng.random(n * n)

    def advance(y0):
        T.set_initial(y0=y0)
        T.step(5e-2)
        return np.array(T.y, copy=True)
    yx, yz = advance(x), advance(z)
    assert abs(math.fsum(yx) - math.fsum(x)) <= 1e-9 * math.fsum(x)
    ymix = advance(0.3 * x - 1.7 * z)
    assert rel_l2(ymix, 0.3 * yx - 1.7 * yz) <= 1e-12
    # the operator itself: constants are in the kernel of L0, interior rows of a linear ramp too
    i, j = np.divmod(np.arange(n * n), n)
    assert np.max(np.abs(T.laplacian(np.full(n * n, 3.25)))) == 0.0
    lap = T.laplacian((2.0 * i + 0.5 * j).astype(float))
    interior = (i > 0) & (i < n - 1) & (j > 0) & (j < n - 1)
    assert np.max(np.abs(lap[interior])) <= 1e-9


def test_hex_2000_curl_of_grad_vanishes_and_advect_is_2_homogeneous():
    """C2 lattice hexagonal_lattice_graph(2000, 2000): B2 B1^T = 0, div(grad) = L0,
    advect(a y) = a^2 advect(y) for a > 0"""
    import gds_b200
    G = gds_b200.NativeGraph('hexagonal', 2000, 2000)
    u, c = gds_b200.edge_gds(G), gds_b200.node_gds(G)
    assert (c.ndim, u.ndim, len(u.faces)) == (8008000, 12007999, 4000000)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(c.ndim)
    g = c.grad(x)
    assert np.max(np.abs(u.curl(g))) <= 1e-13 * np.max(np.abs(g))      # B2 B1^T = 0: only rounding of the 6-term sums is left
    assert rel_l2(u.div(g), c.laplacian(x)) <= 1e-12                   # div = -B1, grad = B1^T, L0 = -B1 B1^T
    y = rng.standard_normal(u.ndim)
    assert rel_l2(u.advect(4.0 * y), 16.0 * u.advect(y)) <= 1e-13
    # the Hodge Laplacian is symmetric: <L a, b> = <a, L b>
    z = rng.standard_normal(u.ndim)
    assert abs(np.dot(u.laplacian(y), z) - np.dot(y, u.laplacian(z))) <= 1e-9 * np.linalg.norm(y) * np.linalg.norm(z)
