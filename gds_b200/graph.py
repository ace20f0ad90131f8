"""Operator assembly: nx.Graph (or a native lattice) -> index maps + device incidence arrays.

Replaces GraphObservable.__init__ / gds.__init__ / the node_gds and edge_gds constructors
(gds/gds.py:16-55,105-109,136-145,194-271).  The index maps follow the reference bit-exactly
(SURVEY.md Appendix C): node index = position in G.nodes(), edge index and stored orientation = position and
(u, v) order in G.edges(), face index = discovery order of the reference's planar face walk with the outer face
removed (gds/utils/graph/generators.py:444-511).  One lowering is shared by every field built on the same graph.
"""
from __future__ import annotations

import ctypes as C
import os

import networkx as nx
import numpy as np

from . import _lib as L
from ._lib import lib, check, ptr
from .device import Context, DeviceGraph
from .utils import bidict


class HostGraph:
    """handle to a native host graph (gds_hostgraph)"""

    def __init__(self, h):
        import weakref
        self.h = h
        self._fin = weakref.finalize(self, lib.gds_hostgraph_destroy, h)
        v, e, f, nnz, ld, hp = L.c_i64(), L.c_i64(), L.c_i64(), L.c_i64(), L.c_i32(), L.c_i32()
        check(lib.gds_hostgraph_sizes(h, C.byref(v), C.byref(e), C.byref(f), C.byref(nnz), C.byref(ld), C.byref(hp)))
        self.V, self.E, self.F, self.face_nnz, self.label_dim, self.has_pos = \
            v.value, e.value, f.value, nnz.value, ld.value, bool(hp.value)

    def edges(self):
        eu, ev = np.empty(self.E, dtype=np.int32), np.empty(self.E, dtype=np.int32)
        check(lib.gds_hostgraph_copy(self.h, None, None, ptr(eu), ptr(ev), None, None))
        return eu, ev

    def labels(self):
        lab = np.empty((self.V, max(1, self.label_dim)), dtype=np.int32)
        if self.label_dim == 0:
            lab[:, 0] = np.arange(self.V)
            return lab
        check(lib.gds_hostgraph_copy(self.h, ptr(lab), None, None, None, None, None))
        return lab

    def pos(self):
        if not self.has_pos:
            return None
        p = np.empty((self.V, 2), dtype=np.float64)
        check(lib.gds_hostgraph_copy(self.h, None, ptr(p), None, None, None, None))
        return p

    def faces(self):
        fp, fn = np.empty(self.F + 1, dtype=np.int64), np.empty(self.face_nnz, dtype=np.int32)
        check(lib.gds_hostgraph_copy(self.h, None, None, None, None, ptr(fp), ptr(fn)))
        return fp, fn


def native_faces(n_nodes, edge_u, edge_v, adj_ptr, adj_nbr, pos):
    """planar face walk on the host (C++), faces as CSR of node indices"""
    h = L.c_vp()
    eu, ev = L.as_i32(edge_u), L.as_i32(edge_v)
    ap, an, p = L.as_i64(adj_ptr), L.as_i32(adj_nbr), L.as_f64(pos)
    check(lib.gds_hostgraph_from_graph(int(n_nodes), eu.size, ptr(eu), ptr(ev), ptr(ap), ptr(an), ptr(p), C.byref(h)))
    return HostGraph(h).faces()


def morton_order(pos: np.ndarray) -> np.ndarray:
    """ordering of points along a Z-order (Morton) curve: nearby points get nearby ranks"""
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((pos - lo) / span * 65535.0).astype(np.uint64), 65535)

    def spread(v):
        v = (v | (v << 8)) & np.uint64(0x00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x33333333)
        v = (v | (v << 1)) & np.uint64(0x55555555)
        return v
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(key, kind='stable')


def locality_permutation(V, edge_u, edge_v, pos, window=1024):
    """Internal node numbering for graphs whose public numbering has no locality (random geometric graphs):
    Z-order of the node positions, then -- inside windows of `window` consecutive nodes -- by decreasing degree, so that
    the 32 rows of a sliced-ELL slice have (nearly) the same width.  Returns i2p (internal -> public) or None when the
    public numbering is already local (lattices).  The public index maps never change; vectors are permuted on the way
    to and from the device, and every row still sums its edges in edge-id order, so results are bit-identical."""
    E = len(edge_u)
    if pos is None or V < 64 or E == 0:
        return None
    if float(np.mean(np.abs(edge_u.astype(np.int64) - edge_v.astype(np.int64)))) <= 0.05 * V:
        return None
    i2p = morton_order(np.asarray(pos, dtype=np.float64))
    window = int(os.environ.get('GDS_B200_SORT_WINDOW', window))     # A/B knob: 0 = keep the pure curve order
    if window <= 1:
        return i2p
    deg = np.bincount(edge_u, minlength=V) + np.bincount(edge_v, minlength=V)
    d = deg[i2p]
    window = min(window, max(32, V // 16))
    win = np.arange(V) // window
    order = np.lexsort((-d, win))            # by window, then by decreasing degree (stable)
    return i2p[order]


def connectivity_order(V, edge_u, edge_v):
    """Ordering for CUTTING a graph that has neither index locality nor node positions into contiguous ranges: reverse
    Cuthill-McKee over the adjacency structure (breadth-first levels, so an index range is a band of a few levels and the
    cut between two ranges is a level boundary).  Returns i2p (position in the order -> public id) or None when the public
    numbering is already local.  Used by `distribute` only (partitioners for graphs that are not lattices);
    graphs with positions take the space-filling curve of `locality_permutation`."""
    E = len(edge_u)
    if V < 64 or E == 0:
        return None
    u, v = np.asarray(edge_u, dtype=np.int64), np.asarray(edge_v, dtype=np.int64)
    if float(np.mean(np.abs(u - v))) <= 0.05 * V:
        return None
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    A = sp.coo_matrix((np.ones(2 * E, dtype=np.int8), (np.concatenate([u, v]), np.concatenate([v, u]))), shape=(V, V)).tocsr()
    return np.ascontiguousarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int64)


class LoweredGraph:
    """Index maps + device operators of one graph, shared by every field built on it.  Use `LoweredGraph.of(G)`.

    Faces are found on first use (the reference walks them eagerly for every field, gds.py:31-36, which is what
    makes node fields on non-planar graphs need the `G.faces = []` hook); the device arrays are lowered when first
    needed and re-lowered once if faces arrive later."""

    def __init__(self, G, ctx: Context = None):
        self.ctx = ctx or Context.default()
        self.G = G
        self.native = isinstance(G, NativeGraph)
        self._dev = None
        self._faces_done = False
        self.outer_face = None
        self.F = 0
        self.i2p = self.p2i = None           # internal <-> public node numbering (None: identical)
        self.part = None                     # Partition when the graph is distributed over several ranks
        dist_spec = G.__dict__.get('_gds_b200_distribute')
        if dist_spec is not None and dist_spec[1] > 1:
            self._init_distributed(G, *dist_spec)
        elif self.native:
            hg = G.hostgraph(with_faces=False)
            self._hg = hg
            self.V, self.E = hg.V, hg.E
            self.edge_u, self.edge_v = hg.edges()
            self.weights = np.ones(self.E) if G.weights is None else L.as_f64(G.weights)
            self.nodes = LazyNodeMap(hg)
            self.edges = LazyEdgeMap(self)
            self._faces = LazyFaceMap(hg, self.nodes)
            if G.kind not in ('triangular', 'hexagonal'):
                self._faces_done = True          # no embedded faces for grids / 3-D grids / geometric graphs
            if G.kind == 'random_geometric':
                self._set_perm(locality_permutation(self.V, self.edge_u, self.edge_v, hg.pos()))
        else:
            self.nodes = {v: i for i, v in enumerate(G.nodes())}                       # gds.py:21
            self.edges = bidict({e: i for i, e in enumerate(G.edges())})               # gds.py:23
            V, E = len(self.nodes), len(self.edges)
            self.V, self.E = V, E
            nodes = self.nodes
            self.edge_u = np.fromiter((nodes[e[0]] for e in self.edges), dtype=np.int32, count=E)
            self.edge_v = np.fromiter((nodes[e[1]] for e in self.edges), dtype=np.int32, count=E)
            # gds.py:39-40, generators.py:600-607: attribute 'weight', default 1.0
            self.weights = np.fromiter((d.get('weight', 1.0) for (_, _, d) in G.edges(data=True)), dtype=np.float64, count=E)
            self._faces = {}
            if hasattr(G, 'faces'):                                                      # gds.py:31-32
                self._set_faces([tuple(f) for f in G.faces])
            pos = nx.get_node_attributes(G, 'pos')
            if len(pos) == V and V:
                P = np.array([pos[v][:2] for v in self.nodes], dtype=np.float64).reshape(V, -1)
                if P.shape[1] == 2:
                    self._set_perm(locality_permutation(V, self.edge_u, self.edge_v, P))

    # -- one rank's part of a distributed graph -------------------------------------------------------------
    def _init_distributed(self, G, rank, world, align, hops=1, blocks=None):
        from .partition import BlockPartition, Partition
        if hops > 1:
            return self._init_distributed_deep(G, rank, world, align, hops)
        if blocks is not None:
            A, B, C = (G.dims[0], G.dims[1], 1) if G.kind == 'grid_2d' else (G.dims[2], G.dims[1], G.dims[0])
            part = BlockPartition(rank, blocks, A, B, C)
            gids = part.global_ids
            if G.kind == 'grid_2d':
                lab = np.stack([gids // B, gids % B], axis=1)
            else:
                lab = np.stack([gids // (B * C), (gids // C) % B, gids % C], axis=1)
            self.nodes = LazyNodeMap(_LabelSource(lab.astype(np.int64), lab.shape[1]))
        elif self.native and G.kind in ('grid_2d', 'grid') and G.weights is None:
            A, B, C = (G.dims[0], G.dims[1], 1) if G.kind == 'grid_2d' else (G.dims[2], G.dims[1], G.dims[0])
            part = Partition.grid_slab(rank, world, A, B, C)      # closed form: only this slab is ever generated
            gids = part.global_ids
            if G.kind == 'grid_2d':
                lab = np.stack([gids // B, gids % B], axis=1)
            else:
                lab = np.stack([gids // (B * C), (gids // C) % B, gids % C], axis=1)
            self.nodes = LazyNodeMap(_LabelSource(lab.astype(np.int64), lab.shape[1]))
        elif self.native:
            hg = G.hostgraph(with_faces=False)
            eu, ev = hg.edges()
            # no locality in the public numbering (random geometric graph): cut along a space-filling curve of the positions
            i2p = locality_permutation(hg.V, eu, ev, hg.pos()) if hg.has_pos else None
            if i2p is not None and not os.environ.get('GDS_B200_NO_RENUMBER'):
                p2i = np.empty(hg.V, dtype=np.int64)
                p2i[i2p] = np.arange(hg.V, dtype=np.int64)
                eu, ev = p2i[eu], p2i[ev]
            else:
                i2p = None
            part = Partition.from_global(rank, world, hg.V, eu, ev, G.weights, align)
            part.set_public_ids(i2p)
            lab = hg.labels()[part.global_ids]
            self.nodes = LazyNodeMap(_LabelSource(lab.astype(np.int64), hg.label_dim))
        else:
            gnodes = list(G.nodes())
            gidx = {v: i for i, v in enumerate(gnodes)}
            E = G.number_of_edges()
            eu = np.fromiter((gidx[e[0]] for e in G.edges()), dtype=np.int64, count=E)
            ev = np.fromiter((gidx[e[1]] for e in G.edges()), dtype=np.int64, count=E)
            w = np.fromiter((d.get('weight', 1.0) for (_, _, d) in G.edges(data=True)), dtype=np.float64, count=E)
            i2p = None
            pos = nx.get_node_attributes(G, 'pos')
            if len(pos) == len(gnodes) and gnodes and not os.environ.get('GDS_B200_NO_RENUMBER'):
                P = np.array([pos[v][:2] for v in gnodes], dtype=np.float64).reshape(len(gnodes), -1)
                if P.shape[1] == 2:
                    i2p = locality_permutation(len(gnodes), eu, ev, P)
            elif not os.environ.get('GDS_B200_NO_RENUMBER'):
                i2p = connectivity_order(len(gnodes), eu, ev)       # no positions: cut along breadth-first levels
            if i2p is not None:
                p2i = np.empty(len(gnodes), dtype=np.int64)
                p2i[i2p] = np.arange(len(gnodes), dtype=np.int64)
                eu, ev = p2i[eu], p2i[ev]
            part = Partition.from_global(rank, world, len(gnodes), eu, ev, w, align)
            part.set_public_ids(i2p)
            self.nodes = {gnodes[int(g)]: i for i, g in enumerate(part.global_ids)}
        self.part = part
        self.V, self.E = part.n_loc, part.E_loc
        self.edge_u, self.edge_v = part.edge_u, part.edge_v
        self.weights = np.ones(self.E) if part.weights is None else part.weights
        self.edges = LazyEdgeMap(self) if self.native else bidict(
            {(a, b): i for i, (a, b) in enumerate(self._local_edge_labels())})
        self._faces = {}
        self._faces_done = True               # node fields only on distributed graphs
        self.native_local = True

    def _init_distributed_deep(self, G, rank, world, align, hops):
        """this rank's part with `hops` ghost layers, its edges and faces (partition.DeepPartition)"""
        from .partition import DeepPartition, split_bounds
        if self.native:
            hg = G.hostgraph(with_faces=G.kind in ('triangular', 'hexagonal'))
            eu, ev = hg.edges()
            w = G.weights
            fp, fn = hg.faces() if hg.F else (None, None)
            glabels = hg.labels()
            Vg = hg.V
        else:
            gnodes = list(G.nodes())
            gidx = {v: i for i, v in enumerate(gnodes)}
            Eg = G.number_of_edges()
            eu = np.fromiter((gidx[e[0]] for e in G.edges()), dtype=np.int64, count=Eg)
            ev = np.fromiter((gidx[e[1]] for e in G.edges()), dtype=np.int64, count=Eg)
            w = np.fromiter((d.get('weight', 1.0) for (_, _, d) in G.edges(data=True)), dtype=np.float64, count=Eg)
            # faces of the WHOLE graph, in the reference's order (gds.py:31-36), found before the cut
            self.nodes, self.edge_u, self.edge_v = gidx, eu.astype(np.int32), ev.astype(np.int32)
            faces = [tuple(f) for f in G.faces] if hasattr(G, 'faces') else self._walk_faces(G)
            fp = np.cumsum([0] + [len(f) for f in faces]).astype(np.int64) if faces else None
            fn = np.fromiter((gidx[n] for f in faces for n in f), dtype=np.int64, count=int(fp[-1])) if faces else None
            Vg = len(gnodes)
        part = DeepPartition(rank, world, Vg, eu, ev, w, split_bounds(Vg, world, align), hops, fp, fn)
        if self.native:
            lab = glabels[part.global_ids]
            self.nodes = LazyNodeMap(_LabelSource(lab.astype(np.int64), hg.label_dim))
        else:
            self.nodes = {gnodes[int(g)]: i for i, g in enumerate(part.global_ids)}
        self.part = part
        self.V, self.E, self.F = part.n_loc, part.E_loc, part.F_loc
        self.edge_u, self.edge_v = part.edge_u, part.edge_v
        self.weights = np.ones(self.E) if part.weights is None else part.weights
        if self.native:
            self.edges = LazyEdgeMap(self)
            self._faces = _LocalFaceMap(self, part)
        else:
            inv = list(self.nodes.keys())
            self.edges = bidict({(inv[u], inv[v]): i for i, (u, v) in enumerate(zip(self.edge_u, self.edge_v))})
            fpl, fnl = part.faces_local
            self._faces = {tuple(inv[j] for j in fnl[fpl[f]:fpl[f + 1]]): f for f in range(part.F_loc)}
        self._faces_done = True
        self.native_local = True

    def _local_edge_labels(self):
        inv = list(self.nodes.keys())
        return [(inv[u], inv[v]) for u, v in zip(self.edge_u, self.edge_v)]

    # -- internal node numbering ---------------------------------------------------------------------------
    def _set_perm(self, i2p):
        if i2p is None or os.environ.get('GDS_B200_NO_RENUMBER'):
            return
        self.i2p = np.ascontiguousarray(i2p, dtype=np.int64)
        self.p2i = np.empty_like(self.i2p)
        self.p2i[self.i2p] = np.arange(self.i2p.size, dtype=np.int64)
        self._dev = None

    def to_dev(self, domain, host_vec):
        """a host vector in public row order -> the order the device arrays use"""
        if domain != L.NODES or self.i2p is None:
            return host_vec
        return np.asarray(host_vec)[self.i2p]

    def from_dev(self, domain, dev_vec):
        if domain != L.NODES or self.p2i is None:
            return dev_vec
        return np.asarray(dev_vec)[self.p2i]

    def rows_to_dev(self, domain, rows):
        """public row indices -> device row indices (same order)"""
        rows = np.asarray(rows, dtype=np.int64)
        if domain != L.NODES or self.p2i is None:
            return rows
        return self.p2i[rows]

    # -- faces -------------------------------------------------------------------------------------------
    def _set_faces(self, faces):
        self._faces = {f: i for i, f in enumerate(faces)}                               # gds.py:35
        self.F = len(self._faces)
        self._face_list = faces
        self._faces_done = True
        self._dev = None                       # lowered again, with B2, on next use

    def ensure_faces(self):
        if self._faces_done:
            return
        if self.native:
            hg = self.G.hostgraph(with_faces=True)
            self._hg = hg
            self._faces = LazyFaceMap(hg, self.nodes)
            self.F = hg.F
            self._faces_done = True
            self._dev = None
        else:
            self._set_faces(self._walk_faces(self.G))                                   # gds.py:34

    @property
    def faces(self):
        self.ensure_faces()
        return self._faces

    def _walk_faces(self, G):
        pos = nx.get_node_attributes(G, 'pos')
        if pos == {}:
            is_planar, _ = nx.check_planarity(G)
            if not is_planar:
                import warnings
                warnings.warn('no faces for non-planar graphs')                     # generators.py:531-534
                return []
            # generators.py:513-518: fixed-seed spring layout, stored on the graph like the reference does
            pos = nx.spring_layout(G, scale=0.9, center=(0, 0), iterations=500, seed=1, dim=2)
            nx.set_node_attributes(G, pos, 'pos')
        nodes = self.nodes
        V = len(nodes)
        P = np.empty((V, 2), dtype=np.float64)
        for v, i in nodes.items():
            P[i, 0], P[i, 1] = pos[v][0], pos[v][1]
        adj_ptr = np.zeros(V + 1, dtype=np.int64)
        adj = []
        for v, i in nodes.items():
            nb = [nodes[u] for u in G.neighbors(v)]
            adj_ptr[i + 1] = adj_ptr[i] + len(nb)
            adj.extend(nb)
        fp, fn = native_faces(V, self.edge_u, self.edge_v, adj_ptr, np.asarray(adj, dtype=np.int32), P)
        inv = list(nodes.keys())
        return [tuple(inv[j] for j in fn[fp[f]:fp[f + 1]]) for f in range(fp.size - 1)]

    # -- device arrays -------------------------------------------------------------------------------------
    @property
    def dev(self) -> DeviceGraph:
        if self._dev is None:
            if self.part is not None:
                unit = bool(np.all(self.weights == 1.0))
                fpl, fnl = self.part.faces_local if self.part.faces_local is not None and self.F else (None, None)
                self._dev = DeviceGraph(self.ctx, self.V, self.edge_u, self.edge_v, None if unit else self.weights, fpl, fnl)
            elif self.native and self.p2i is None:
                self._dev = DeviceGraph(self.ctx, self.V, None, None, self.G.weights, hostgraph=self._hg.h)
            elif self.native:
                fp, fn = (self._hg.faces() if self._hg.F else (None, None))
                self._dev = DeviceGraph(self.ctx, self.V, self.p2i[self.edge_u], self.p2i[self.edge_v], self.G.weights,
                                        fp, None if fn is None else self.p2i[fn])
            else:
                face_ptr = face_nodes = None
                if self.F:
                    faces, nodes = self._face_list, self.nodes
                    face_ptr = np.cumsum([0] + [len(f) for f in faces]).astype(np.int64)
                    face_nodes = np.fromiter((nodes[n] for f in faces for n in f), dtype=np.int32, count=int(face_ptr[-1]))
                unit = bool(np.all(self.weights == 1.0))
                eu, ev = self.edge_u, self.edge_v
                if self.p2i is not None:
                    eu, ev = self.p2i[eu], self.p2i[ev]
                    face_nodes = None if face_nodes is None else self.p2i[face_nodes]
                self._dev = DeviceGraph(self.ctx, self.V, eu, ev, None if unit else self.weights, face_ptr, face_nodes)
        return self._dev

    def component_labels(self):
        """connected component of every node, in DEVICE row order (cached); used to keep right-hand sides of the
        singular graph Laplacian in its range"""
        if getattr(self, '_components', None) is None:
            import scipy.sparse as sp
            from scipy.sparse.csgraph import connected_components
            A = sp.coo_matrix((np.ones(self.E, dtype=np.int8), (self.edge_u, self.edge_v)), shape=(self.V, self.V))
            ncomp, lab = connected_components(A, directed=False)
            self._components = (ncomp, np.ascontiguousarray(self.to_dev(L.NODES, lab)))
        return self._components

    def rows(self, domain):
        if domain == L.FACES:
            self.ensure_faces()
        return {L.NODES: self.V, L.EDGES: self.E, L.FACES: self.F}[domain]

    @staticmethod
    def _signature(G):
        """cheap fingerprint of what the lowering depends on: a graph edited after a field was built is lowered again"""
        if isinstance(G, NativeGraph):
            return (G.kind, G.dims, id(G.weights), G.__dict__.get('_gds_b200_distribute'))
        wsum = sum(d.get('weight', 1.0) for (_, _, d) in G.edges(data=True))
        nf = len(G.faces) if hasattr(G, 'faces') else -1
        return (G.number_of_nodes(), G.number_of_edges(), wsum, nf, G.__dict__.get('_gds_b200_distribute'))

    @staticmethod
    def of(G, need_faces: bool = False) -> 'LoweredGraph':
        low = G.__dict__.get('_gds_b200_lowered')
        sig = LoweredGraph._signature(G)
        if low is not None and low._sig != sig:
            low = None
        if low is None:
            low = LoweredGraph(G)
            low._sig = sig
            G.__dict__['_gds_b200_lowered'] = low
        if need_faces:
            low.ensure_faces()
        return low


# ------------------------------------------------------------------------------------------------
# native lattices (10^8 .. 10^9 edges: an nx.Graph of that size does not fit in memory)
# ------------------------------------------------------------------------------------------------
class NativeGraph:
    """A graph generated natively with networkx-identical node / edge / face order.  Stands in for nx.Graph in the
    field constructors.  kind: 'grid_2d' (m, n), 'grid' (d0, d1, d2), 'triangular' (m, n), 'hexagonal' (m, n),
    'random_geometric' (n, radius, seed)."""
    KINDS = {'grid_2d': 0, 'grid': 1, 'triangular': 2, 'hexagonal': 3, 'random_geometric': 4}

    def __init__(self, kind, *dims, weights=None):
        self.kind, self.dims, self.weights = kind, tuple(dims), weights
        self._hg = {}

    def hostgraph(self, with_faces=False) -> HostGraph:
        with_faces = bool(with_faces) and self.kind in ('triangular', 'hexagonal')
        hg = self._hg.get(True) or (self._hg.get(False) if not with_faces else None)
        if hg is None:
            h = L.c_vp()
            k = self.KINDS[self.kind]
            if self.kind == 'random_geometric':
                n, r, seed = self.dims
                check(lib.gds_hostgraph_lattice(k, int(n), int(seed), 0, float(r), 0, C.byref(h)))
            else:
                d = list(self.dims) + [0, 0]
                check(lib.gds_hostgraph_lattice(k, int(d[0]), int(d[1]), int(d[2]), 0.0, int(with_faces), C.byref(h)))
            hg = HostGraph(h)
            self._hg[with_faces] = hg
        return hg

    def number_of_nodes(self):
        return self.hostgraph().V

    def number_of_edges(self):
        return self.hostgraph().E


class _LabelSource:
    """label array standing in for a host graph (rows of a distributed graph)"""

    def __init__(self, labels, label_dim):
        self._labels, self.V, self.label_dim = labels, int(labels.shape[0]), int(label_dim)

    def labels(self):
        return self._labels


def distribute(G, rank=None, world=None, align=1, hops=1, blocks=None):
    """Mark graph G (nx.Graph or NativeGraph) as partitioned over `world` ranks; fields built on it afterwards hold
    this rank's rows and exchange boundary values every stage.  Defaults come from the torch.distributed / torchrun
    environment.  hops = 1: node fields with one-hop laws (owned nodes, then one layer of ghosts).  hops > 1: `hops`
    layers of ghost nodes with every edge and face inside them -- edge and face fields, and laws that chain gathers
    (Hodge Laplacian, self-advection): intermediates are recomputed on the ghost layers, only the state is exchanged."""
    import os
    rank = int(os.environ.get('RANK', '0')) if rank is None else int(rank)
    world = int(os.environ.get('WORLD_SIZE', '1')) if world is None else int(world)
    if blocks is not None:
        # block decomposition of a native grid: (pa, pb, pc) blocks along the axes of node labels (a, b, c)
        blocks = tuple(int(x) for x in blocks) + (1,) * (3 - len(blocks))
        assert int(np.prod(blocks)) == world, 'the blocks must multiply to the number of ranks'
        assert isinstance(G, NativeGraph) and G.kind in ('grid_2d', 'grid') and G.weights is None and hops == 1, \
            'block decomposition is available for unweighted native grids'
    G.__dict__['_gds_b200_distribute'] = (rank, world, int(align), int(hops)) + ((blocks,) if blocks is not None else ())
    G.__dict__.pop('_gds_b200_lowered', None)
    return G


class _LocalFaceMap:
    """faces of one rank of a distributed native graph, as tuples of node labels -> local face index (lazy)"""

    def __init__(self, low, part):
        self._low, self._part, self._dict = low, part, None

    def _d(self):
        if self._dict is None:
            fp, fn = self._part.faces_local
            lab = self._low.nodes.labels()
            key = self._low.nodes._key
            self._dict = {tuple(key(lab[j]) for j in fn[fp[f]:fp[f + 1]]): f for f in range(fp.size - 1)}
        return self._dict

    def __len__(self): return self._part.F_loc
    def __getitem__(self, k): return self._d()[k]
    def __contains__(self, k): return k in self._d()
    def __iter__(self): return iter(self._d())
    def keys(self): return self._d().keys()
    def items(self): return self._d().items()
    def values(self): return self._d().values()


class LazyNodeMap:
    """label -> index map of a native graph; materialised only on demand."""

    def __init__(self, hg: HostGraph):
        self._hg = hg
        self._labels = None
        self._dict = None

    def labels(self):
        if self._labels is None:
            self._labels = self._hg.labels()
        return self._labels

    def _key(self, row):
        return int(row[0]) if self._hg.label_dim == 0 else tuple(int(x) for x in row)

    def _d(self):
        if self._dict is None:
            self._dict = {self._key(r): i for i, r in enumerate(self.labels())}
        return self._dict

    def __len__(self): return self._hg.V
    def __getitem__(self, k): return self._d()[k]
    def __contains__(self, k): return k in self._d()
    def __iter__(self): return iter(self._d())
    def keys(self): return self._d().keys()
    def items(self): return self._d().items()
    def values(self): return self._d().values()


class LazyEdgeMap:
    def __init__(self, low):
        self._low = low
        self._dict = None

    def _d(self):
        if self._dict is None:
            lab = self._low.nodes.labels()
            key = self._low.nodes._key
            self._dict = bidict({(key(lab[u]), key(lab[v])): i
                                 for i, (u, v) in enumerate(zip(self._low.edge_u, self._low.edge_v))})
        return self._dict

    def __len__(self): return self._low.E
    def __getitem__(self, k): return self._d()[k]
    def __contains__(self, k): return k in self._d()
    def __iter__(self): return iter(self._d())
    def keys(self): return self._d().keys()
    def items(self): return self._d().items()
    def values(self): return self._d().values()


class LazyFaceMap:
    def __init__(self, hg, nodes):
        self._hg, self._nodes, self._dict = hg, nodes, None

    def _d(self):
        if self._dict is None:
            fp, fn = self._hg.faces()
            lab = self._nodes.labels()
            key = self._nodes._key
            self._dict = {tuple(key(lab[j]) for j in fn[fp[f]:fp[f + 1]]): f for f in range(fp.size - 1)}
        return self._dict

    def __len__(self): return self._hg.F
    def __getitem__(self, k): return self._d()[k]
    def __contains__(self, k): return k in self._d()
    def __iter__(self): return iter(self._d())
    def keys(self): return self._d().keys()
    def items(self): return self._d().items()
    def values(self): return self._d().values()
