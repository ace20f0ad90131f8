"""Graphs whose 16-bit relative neighbour indices need the escape value (-32768: take this quad's indices from the int32
array) -- a ring (closing edge 0 <-> n-1), a grid with one long-range edge, and the local graph of a slab behind the
first, whose ghost lattice row is numbered AFTER the owned rows so that a whole row of neighbours is > 32767 rows away
(escapes that are aligned runs) -- against a scipy Laplacian: the operator and 3 RK4 steps."""
import networkx as nx
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grid_first_row_last(n=200):
    H = nx.grid_2d_graph(n, n)
    G = nx.Graph()
    G.add_nodes_from([v for v in H.nodes() if v[0] > 0] + [v for v in H.nodes() if v[0] == 0])
    G.add_edges_from(H.edges())
    return G


def _grid_long_edge(n=200):
    G = nx.grid_2d_graph(n, n)
    G.add_edge((0, 0), (n - 1, n - 1))
    return G


GRAPHS = {'cycle_40000': lambda: nx.cycle_graph(40000), 'grid_200x200_plus_long_edge': _grid_long_edge,
          'grid_200x200_first_row_last': _grid_first_row_last}


@pytest.mark.parametrize('name', list(GRAPHS))
def test_index_escape_paths_match_scipy(name):
    import gds_b200
    G = GRAPHS[name]()
    G.faces = []
    T = gds_b200.node_gds(G)
    L0 = -nx.laplacian_matrix(G, nodelist=list(G.nodes())).astype(float).tocsr()
    y0 = np.random.default_rng(1).standard_normal(T.ndim)
    assert np.linalg.norm(np.asarray(T.laplacian(y0)) - L0 @ y0) <= 1e-12 * np.linalg.norm(L0 @ y0)
    h = 1e-2
    T.set_evolution(dydt=lambda t, y: 0.5 * T.laplacian(y), max_step=h)
    T.set_initial(y0=y0)
    y = y0.copy()
    f = lambda v: 0.5 * (L0 @ v)
    for _ in range(3):
        k1 = f(y); k2 = f(y + h / 2 * k1); k3 = f(y + h / 2 * k2); k4 = f(y + h * k3)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        T.step(h)
    assert np.linalg.norm(np.asarray(T.y) - y) <= 1e-12 * np.linalg.norm(y)
