"""Graph-construction helpers the reference's examples build their domains with (gds/utils/graph/generators.py): planar
lattices with their boundary sub-graphs, edge weights, and the embedded faces -- the inputs of the stepping path.
Pure host code over networkx; node order, G.edges() order, positions and boundary sets are those of the reference
(tests/test_golden_fresh_cpu.py compares them with the unmodified reference object by object).

Not provided: the renderer-oriented pieces of that module (lattice_components, periodic variants sanitised for Bokeh,
3-D meshes, Voronoi lattices) -- they do not feed the stepping path.
"""
from __future__ import annotations

import math

import networkx as nx
import numpy as np

__all__ = ['set_grid_graph_layout', 'square_lattice', 'triangular_lattice', 'hexagonal_lattice', 'get_planar_boundary',
           'embedded_faces', 'get_edge_weights', 'set_edge_weights', 'get_edge_lengths', 'edge_domain', 'remove_edges',
           'remove_pos', 'degree_matrix']


def set_grid_graph_layout(G: nx.Graph):
    """'pos' for a grid_2d_graph with nodes (i, j): the longer side spans [-1, 1) (generators.py:24-40)"""
    n = max(v[0] for v in G.nodes()) + 1
    m = max(v[1] for v in G.nodes()) + 1
    big = max(m, n)
    dh, x0, y0 = 1 / big, -n / big, -m / big
    present = set(G.nodes())
    pos = {(i, j): np.array([2 * i * dh + x0, 2 * j * dh + y0]) for i in range(n) for j in range(m) if (i, j) in present}
    nx.set_node_attributes(G, pos, 'pos')


def _side(G, nodes):
    return G.subgraph([v for v in nodes]).copy()


def _reject(**flags):
    for k, v in flags.items():
        if v:
            raise NotImplementedError(f'{k} is a renderer-side feature of the reference and is not provided')


def square_lattice(m: int, n: int, diagonals=False, with_boundaries=False, with_lattice_components=False, **kwargs):
    """n x m square grid (nodes (i, j), i < n, j < m) with positions; with_boundaries also returns the left, right, top
    and bottom sides as sub-graphs (generators.py:42-66)"""
    _reject(with_lattice_components=with_lattice_components)
    G = nx.grid_2d_graph(n, m, **kwargs)
    set_grid_graph_layout(G)
    if diagonals:
        for i in range(n - 1):
            for j in range(m - 1):
                G.add_edge((i, j), (i + 1, j + 1))
                G.add_edge((i, j + 1), (i + 1, j))
    if not with_boundaries:
        return G
    left, right = [(0, j) for j in range(m)], [(n - 1, j) for j in range(m)]
    top, bottom = [(i, m - 1) for i in range(n)], [(i, 0) for i in range(n)]
    return G, (_side(G, left), _side(G, right), _side(G, top), _side(G, bottom))


def triangular_lattice(m, n, with_boundaries=False, with_lattice_components=False, **kwargs):
    """networkx's triangular lattice with the reference's boundary sets (generators.py:105-143): the right side is the
    zig-zag column the generator ends on"""
    _reject(with_lattice_components=with_lattice_components, periodic='periodic' in kwargs)
    G = nx.triangular_lattice_graph(m, n, **kwargs)
    if not with_boundaries:
        return G
    half = n // 2
    right = [(half, 2 * i + 1) for i in range(m // 2 + 1)]
    right += [(half + 1, i) for i in range(m + 1)] if n % 2 == 1 else [(half, 2 * i) for i in range(m // 2 + 1)]
    return G, (_side(G, [(0, i) for i in range(m + 1)]), _side(G, [v for v in right if v in G]),
               _side(G, [(j, m) for j in range(n)]), _side(G, [(j, 0) for j in range(n)]))


def hexagonal_lattice(m, n, with_boundaries=False, with_lattice_components=False, with_faces=True, **kwargs):
    """networkx's hexagonal lattice; G.faces holds the hexagons in the reference's order (generators.py:165-200)"""
    _reject(with_lattice_components=with_lattice_components, periodic='periodic' in kwargs)
    G = nx.hexagonal_lattice_graph(m, n, **kwargs)
    if with_faces:
        G.faces, _ = embedded_faces(G)
    if not with_boundaries:
        return G
    rows = range(2 * m + 2)
    top = [(j, 2 * m) for j in range(n + 1)] + [(j, 2 * m + 1) for j in range(n + 1)]
    bottom = [(j, 0) for j in range(n + 1)] + [(j, 1) for j in range(n + 1)]
    return G, (_side(G, [(0, i) for i in rows]), _side(G, [(n, i) for i in rows]), _side(G, top), _side(G, bottom))


def get_planar_boundary(G: nx.Graph):
    """(boundary, left, right, top, bottom) of a positioned planar graph: the extreme node of every horizontal /
    vertical line of nodes, with the edges of G between them in their stored orientation (generators.py:396-429)"""
    pos = nx.get_node_attributes(G, 'pos')
    nodes = set(G.nodes())
    lo_x, hi_x, lo_y, hi_y = {}, {}, {}, {}
    for v in nodes:
        x, y = pos[v]
        lo_x[y] = min(lo_x.get(y, x), x)
        hi_x[y] = max(hi_x.get(y, x), x)
        lo_y[x] = min(lo_y.get(x, y), y)
        hi_y[x] = max(hi_y.get(x, y), y)
    whole, left, right, top, bottom = (nx.Graph() for _ in range(5))
    for v in nodes:
        x, y = pos[v]
        for hit, side in ((x == lo_x[y], left), (x == hi_x[y], right), (y == lo_y[x], bottom), (y == hi_y[x], top)):
            if hit:
                side.add_node(v)
                whole.add_node(v)
    stored = set(G.edges())
    for side in (whole, left, right, top, bottom):
        members = list(side.nodes())
        for a in members:
            for b in members:
                if (a, b) in stored:
                    side.add_edge(a, b)
                elif (b, a) in stored:
                    side.add_edge(b, a)
    return whole, left, right, top, bottom


def _turn(pos, a, b, c):
    """signed angle in (-pi, pi] from the direction a -> b to the direction b -> c"""
    t = (np.arctan2(pos[c][1] - pos[b][1], pos[c][0] - pos[b][0])
         - np.arctan2(pos[b][1] - pos[a][1], pos[b][0] - pos[a][0]))
    t = np.abs(t) % (2 * np.pi) if t >= 0 else 2 * np.pi - np.abs(t) % (2 * np.pi)
    return t - 2 * np.pi if t > np.pi else t


def embedded_faces(G: nx.Graph, use_spring_default=True):
    """(faces, outer_face) of a planar graph, generators.py:444-534: counter-clockwise node tuples in the discovery
    order of the half-edge walk over G.edges(), the unbounded face split off.  The inner faces come from the library's
    native walk (the one every edge / face field of G uses); the outer face is the walk from the first half-edge they
    leave uncovered."""
    if not use_spring_default and nx.get_node_attributes(G, 'pos') == {}:
        raise NotImplementedError('combinatorial embeddings (use_spring_default=False) are not provided')
    from .graph import LoweredGraph
    faces = list(LoweredGraph.of(G).faces.keys())
    pos = nx.get_node_attributes(G, 'pos')
    if not faces and pos == {}:
        return [], None                      # non-planar: the reference warns and gives up the same way
    covered = set()
    for f in faces:
        covered.update(zip(f[-1:] + f[:-1], f))
    outer = None
    for u, v in G.edges():
        for tail, head in ((u, v), (v, u)):
            if outer is not None or (tail, head) in covered:
                continue
            path = [head]
            while True:
                best, best_angle = None, -math.inf
                for w in G.neighbors(head):
                    if w != tail:
                        ang = _turn(pos, tail, head, w)
                        if ang > best_angle:
                            best, best_angle = w, ang
                if best is None:             # a dangling edge: the reference drops this walk
                    path = None
                    break
                if best == path[0]:
                    break
                path.append(best)
                tail, head = head, best
            if path is not None:
                outer = tuple(path)
    return faces, outer


def get_edge_lengths(G: nx.Graph):
    pos = nx.get_node_attributes(G, 'pos')
    assert pos != {}, 'need a node embedding'
    return {(u, v): np.sqrt((pos[u][0] - pos[v][0]) ** 2 + (pos[u][1] - pos[v][1]) ** 2) for u, v in G.edges()}


def get_edge_weights(G: nx.Graph, key='weight', default=1.0):
    return {(u, v): d.get(key, default) for u, v, d in G.edges(data=True)}


def set_edge_weights(G: nx.Graph, w, key='weight'):
    """w: a number or a function of the edge (u, v) (generators.py:609-617)"""
    if not callable(w):
        value = float(w)
        w = lambda e: value   # noqa: E731
    for u, v in G.edges():
        G[u][v][key] = w((u, v))
    return G


def edge_domain(G, nodes):
    """edges of G with both endpoints in `nodes`, stored orientation (generators.py:708-712)"""
    nodes = set(nodes)
    return ((u, v) for u, v in G.edges() if u in nodes and v in nodes)


def remove_edges(G: nx.Graph, edges):
    for u, v in edges:
        if G.has_edge(u, v):
            G.remove_edge(u, v)
    return G


def remove_pos(G):
    for v in G.nodes():
        G.nodes[v].pop('pos', None)
    return G


def degree_matrix(G):
    """diagonal matrix of the weighted degrees, sum over incident edges of their 'weight' (gds/utils/graph/common.py:5-10)"""
    import scipy.sparse as sp
    index = {v: i for i, v in enumerate(G.nodes())}
    deg = np.zeros(len(index))
    for u, v, d in G.edges(data=True):
        w = d.get('weight', 1)
        deg[index[u]] += w
        if u != v:
            deg[index[v]] += w
    return sp.diags(deg, format='csr')
