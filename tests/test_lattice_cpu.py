"""Native graph generators (C++, behind the C-ABI) against networkx: node order, G.edges() order and orientation,
positions, and the planar face walk against the oracle's restatement of gds/utils/graph/generators.py:444-511
(SURVEY.md Appendix C gate: element-wise equality)."""
import math

import networkx as nx
import numpy as np
import pytest

import gds_b200
from gds_b200.graph import NativeGraph
from oracle import gds_oracle


def check_against_nx(ng: NativeGraph, G, faces=False):
    hg = ng.hostgraph(with_faces=faces)
    nodes = list(G.nodes())
    assert hg.V == len(nodes) and hg.E == G.number_of_edges()
    lab = hg.labels()
    if hg.label_dim:
        assert [tuple(int(x) for x in r) for r in lab] == [tuple(n) for n in nodes]
    else:
        assert [int(r[0]) for r in lab] == nodes
    idx = {n: i for i, n in enumerate(nodes)}
    eu, ev = hg.edges()
    want = np.array([[idx[a], idx[b]] for a, b in G.edges()], dtype=np.int64).reshape(-1, 2)
    assert np.array_equal(np.stack([eu, ev], axis=1), want)
    pos = nx.get_node_attributes(G, 'pos')
    if pos:
        P = hg.pos()
        assert np.array_equal(P, np.array([pos[n] for n in nodes], dtype=float))
    if faces:
        ref_faces, _ = gds_oracle.embedded_faces(G)
        fp, fn = hg.faces()
        got = [tuple(nodes[j] for j in fn[fp[f]:fp[f + 1]]) for f in range(fp.size - 1)]
        assert got == [tuple(f) for f in ref_faces]


@pytest.mark.parametrize('m,n', [(1, 1), (2, 5), (7, 3), (10, 10)])
def test_grid_2d(m, n):
    check_against_nx(NativeGraph('grid_2d', m, n), nx.grid_2d_graph(m, n))


@pytest.mark.parametrize('dims', [(2, 2, 2), (3, 4, 5), (6, 2, 3), (1, 4, 3)])
def test_grid_3d(dims):
    check_against_nx(NativeGraph('grid', *dims), nx.grid_graph(dim=list(dims)))


@pytest.mark.parametrize('m,n', [(1, 1), (2, 3), (4, 6), (5, 7), (6, 9), (3, 12)])
def test_triangular(m, n):
    check_against_nx(NativeGraph('triangular', m, n), nx.triangular_lattice_graph(m, n), faces=True)


@pytest.mark.parametrize('m,n', [(1, 1), (2, 2), (3, 4), (4, 5), (6, 8), (5, 2)])
def test_hexagonal(m, n):
    check_against_nx(NativeGraph('hexagonal', m, n), nx.hexagonal_lattice_graph(m, n), faces=True)


@pytest.mark.parametrize('n,seed', [(60, 3), (400, 42), (1000, 7)])
def test_random_geometric(n, seed):
    r = math.sqrt(8 / (math.pi * n))
    check_against_nx(NativeGraph('random_geometric', n, r, seed), nx.random_geometric_graph(n, r, seed=seed))


def test_closed_form_sizes():
    # SURVEY.md 8a closed forms, used to size the BASELINE configs without building them
    hg = NativeGraph('hexagonal', 20, 30).hostgraph(with_faces=True)
    m, n = 20, 30
    assert (hg.V, hg.E, hg.F) == (2 * (m + 1) * (n + 1) - 2, 3 * m * n + 2 * (m + n) - 1, m * n)
    hg = NativeGraph('triangular', 20, 30).hostgraph(with_faces=True)
    V = (m + 1) * (n // 2 + 1)
    assert (hg.V, hg.F, hg.E) == (V, m * n, V + m * n - 1)
    hg = NativeGraph('grid', 7, 7, 7).hostgraph()
    assert hg.E == 3 * 7 * 7 * 6


def test_native_graph_field_maps_are_lazy_but_complete():
    import os
    os.environ['GDS_B200_DEVICE'] = '-1'
    gds_b200.Context._default.clear()
    try:
        T = gds_b200.node_gds(NativeGraph('grid_2d', 4, 5))
        G = nx.grid_2d_graph(4, 5)
        assert list(T.X.keys()) == list(G.nodes())
        assert T.ndim == 20
        e = gds_b200.edge_gds(NativeGraph('hexagonal', 2, 3))
        H = nx.hexagonal_lattice_graph(2, 3)
        assert list(e.X.keys()) == list(H.edges())
        assert e.X[list(H.edges())[3][::-1]] == 3      # orientation-agnostic lookup (bidict)
    finally:
        del os.environ['GDS_B200_DEVICE']
        gds_b200.Context._default.clear()
