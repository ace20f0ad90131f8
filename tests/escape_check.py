"""GPU spot check (not collected by pytest): graphs whose 16-bit relative neighbour indices need the escape value -- a
ring (closing edge 0 <-> n-1) and a grid with one long-range edge -- against a scipy Laplacian, operator and RK4 steps.
    python tests/escape_check.py"""
import os
import sys

import networkx as nx
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check(name, G):
    import gds_b200
    G.faces = []
    T = gds_b200.node_gds(G)
    L0 = -nx.laplacian_matrix(G, nodelist=list(G.nodes())).astype(float).tocsr()
    y0 = np.random.default_rng(1).standard_normal(T.ndim)
    e_op = np.linalg.norm(np.asarray(T.laplacian(y0)) - L0 @ y0) / np.linalg.norm(L0 @ y0)
    h = 1e-2
    T.set_evolution(dydt=lambda t, y: 0.5 * T.laplacian(y), max_step=h)
    T.set_initial(y0=y0)
    y = y0.copy()
    f = lambda v: 0.5 * (L0 @ v)
    for _ in range(3):
        k1 = f(y); k2 = f(y + h / 2 * k1); k3 = f(y + h / 2 * k2); k4 = f(y + h * k3)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        T.step(h)
    e_tr = np.linalg.norm(np.asarray(T.y) - y) / np.linalg.norm(y)
    ok = e_op < 1e-12 and e_tr < 1e-12
    print(f'{name}: operator {e_op:.2e}, 3 RK4 steps {e_tr:.2e} -> {"ok" if ok else "FAIL"}', flush=True)
    return ok


if __name__ == '__main__':
    G2 = nx.grid_2d_graph(200, 200)
    G2.add_edge((0, 0), (199, 199))
    # the local graph of a slab behind the first: its ghost lattice row is numbered AFTER the owned rows, so a whole row of
    # neighbours is > 32767 rows away (escapes that are aligned runs) -- here: a 200 x 200 grid whose first row comes last
    H = nx.grid_2d_graph(200, 200)
    G3 = nx.Graph()
    G3.add_nodes_from([n for n in H.nodes() if n[0] > 0] + [n for n in H.nodes() if n[0] == 0])
    G3.add_edges_from(H.edges())
    good = (check('cycle_40000', nx.cycle_graph(40000)) & check('grid_200x200_plus_long_edge', G2)
            & check('grid_200x200_first_row_last', G3))
    sys.exit(0 if good else 1)
