"""Parity at the sizes bench.py runs (BASELINE configurations at full size), not only through size-independent
properties: the CUDA path against oracle/array_oracle.py -- the array-form CPU oracle, pinned against gds_oracle.py and
through it against the reference's golden vectors in tests/test_array_oracle_cpu.py -- after 3 RK4 steps, relative
L2 <= 1e-10 (the north-star tolerance).  Both sides get the same inputs: the lattice's edge list in G.edges() order (the
native generators are checked element-wise against networkx in tests/test_lattice_cpu.py), the weights, the faces, the
boundary rows and the initial arrays.  These are the exact instantiations the bench times: 16-bit relative indices + run
gathers + computed slice positions (unit weights), the weighted quad kernel, and the term-list kernels of config 2."""
import math

import numpy as np
import pytest

import bench
from adaptors import product_api, rel_l2
from oracle import array_oracle as ao

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _host_arrays(G, with_faces=False):
    hg = G.hostgraph(with_faces=with_faces)
    eu, ev = hg.edges()
    return hg, eu.astype(np.int64), ev.astype(np.int64)


@pytest.mark.parametrize('weighted', [False, True], ids=['unit', 'weighted'])
def test_heat_7072_three_steps_against_the_oracle(weighted):
    n, nsteps = 7072, 3
    prob = bench.heat2d_problem(product_api(), n, native=True, weighted=weighted)
    T, h = prob['stepper'], prob['h']
    y0 = np.array(T.y, copy=True)
    eng = T._ensure_engine()
    eng.run(eng.plan_steps(nsteps, h))
    got = np.array(T.y, copy=True)
    # the oracle side, from the same inputs
    hg, eu, ev = _host_arrays(T.G)
    w = None if not weighted else np.asarray(T.G.weights, dtype=float)
    B = ao.incidence(hg.V, eu, ev, w)
    jj = np.arange(n)
    D = np.concatenate([jj, (n - 1) * n + jj])
    g = lambda t: np.concatenate([np.sin(t + jj / 4) ** 2, np.zeros(n)])
    lap = ao.node_laplacian(B, D)
    want = ao.step_fixed(lambda t, y, ya: lap(y), y0, 0., h, nsteps, D, g)
    assert np.all(np.isfinite(got))
    assert rel_l2(got, want) <= TOL
    # the change over the steps, too: a kernel that left the state untouched would pass the line above on a slowly
    # varying field
    assert rel_l2(got - y0, want - y0) <= 1e-8


def test_hex_2000_advection_diffusion_three_steps_against_the_oracle():
    n, nsteps = 2000, 3
    nu, kappa = 0.1, 0.05
    prob = bench.hex_advdiff_problem(product_api(), n, native=True)
    u, c = prob['fields']
    st, h = prob['stepper'], prob['h']
    u0, c0 = np.array(u.y, copy=True), np.array(c.y, copy=True)
    eng = st._ensure_engine()
    eng.run(eng.plan_steps(nsteps, h))
    got_u, got_c = np.array(u.y, copy=True), np.array(c.y, copy=True)
    hg, eu, ev = _host_arrays(u.G, with_faces=True)
    fp, fn = hg.faces()
    V, E = hg.V, hg.E
    B = ao.incidence(V, eu, ev, None)
    B2 = ao.curl_matrix(V, eu, ev, None, fp, fn)
    pos = hg.pos()
    hh = math.sqrt(3) / 2
    lo, hi = pos[:, 1].min(), pos[:, 1].max()
    bnd = (pos[:, 1] < lo + 0.6 * hh) | (pos[:, 1] > hi - 0.6 * hh)
    D_e = np.nonzero(bnd[eu] & bnd[ev])[0]
    N_n = np.nonzero(pos[:, 1] > hi - 1e-9)[0]
    elap = ao.edge_laplacian(B, B2, D_e)
    nlap = ao.node_laplacian(B, (), N_n, -0.1)
    adv = ao.node_advect(B)

    def rhs(t, y, ya):
        # `c.advect(u)` reads the ACCEPTED state of both fields (no y argument: gds.py:166-177 with y = self.y)
        return np.concatenate([nu * elap(y[:E]), -adv(ya[:E], ya[E:]) + kappa * nlap(y[E:])])
    y0 = np.concatenate([u0, c0])
    want = ao.step_fixed(rhs, y0, 0., h, nsteps, D_e, lambda t: np.zeros(D_e.size))
    assert rel_l2(got_u, want[:E]) <= TOL and rel_l2(got_c, want[E:]) <= TOL
    assert rel_l2(got_u - u0, want[:E] - u0) <= 1e-8 and rel_l2(got_c - c0, want[E:] - c0) <= 1e-8
