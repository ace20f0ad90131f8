"""Graph partitioning for multi-GPU stepping -- no counterpart in the single-process reference.

Nodes are split into contiguous index ranges (slabs of a row-major lattice; `align` keeps cuts on plane boundaries).
Graphs whose public numbering has no locality (random geometric graphs) are first ordered along a space-filling curve
over the node positions, so that a contiguous range is a spatially compact part (the coordinate-bisection family of
partitions; `Partition.set_public_ids`); the public ids never change.
Rank r lowers its LOCAL graph: owned nodes [lo, hi) first, then the one-hop ghost nodes sorted by global id (hence
grouped by owner), with every edge that touches an owned node, kept in GLOBAL edge order -- an owned row therefore sums
its incident edges in the same order as on one GPU and results do not depend on the number of parts.

Per stage each rank sends, to every neighbour q, its owned rows adjacent to q's range (sorted by global id) and
receives q's into the ghost segment; both sides derive these lists from their own cut edges, no set-up exchange is
needed.
"""
from __future__ import annotations

import numpy as np


def split_bounds(V: int, world: int, align: int = 1) -> np.ndarray:
    """world+1 cut points of [0, V) in multiples of `align` (the last one is V)"""
    units = -(-V // align)
    b = np.array([min(V, (units * k // world) * align) for k in range(world + 1)], dtype=np.int64)
    b[-1] = V
    return b


class Partition:
    """local graph of one rank"""
    pub = p2i = None      # renumbered cut order <-> public ids (set_public_ids); None: ranges of the public numbering

    def __init__(self, rank, world, V_global, bounds, eu_g, ev_g, weights, ghosts=None):
        """eu_g / ev_g: GLOBAL endpoint ids of the local edges (every edge touching [lo, hi)), in global edge order"""
        self.rank, self.world, self.V_global = int(rank), int(world), int(V_global)
        self.bounds = np.asarray(bounds, dtype=np.int64)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi, self.n_own = lo, hi, hi - lo
        eu_g = np.asarray(eu_g, dtype=np.int64)
        ev_g = np.asarray(ev_g, dtype=np.int64)
        out_u = (eu_g < lo) | (eu_g >= hi)
        out_v = (ev_g < lo) | (ev_g >= hi)
        assert not np.any(out_u & out_v), 'local edge list holds an edge without an owned endpoint'
        if ghosts is None:
            ghosts = np.unique(np.concatenate([eu_g[out_u], ev_g[out_v]]))
        self.ghosts = np.asarray(ghosts, dtype=np.int64)                 # sorted global ids
        self.n_ghost = int(self.ghosts.size)
        self.n_loc = self.n_own + self.n_ghost

        def to_local(g, outside):
            loc = g - lo
            if outside.any():
                loc = loc.copy()
                loc[outside] = self.n_own + np.searchsorted(self.ghosts, g[outside])
            return loc.astype(np.int32)
        self.edge_u = to_local(eu_g, out_u)
        self.edge_v = to_local(ev_g, out_v)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.E_loc = int(self.edge_u.size)
        # ghost segment by owner, and the owned rows each neighbour needs (owned endpoints of the cut edges)
        owner = np.searchsorted(self.bounds, self.ghosts, side='right') - 1
        self.recv_counts = np.bincount(owner, minlength=world).astype(np.int64)
        self.recv_offsets = np.concatenate([[0], np.cumsum(self.recv_counts)])[:-1]
        cut_owned = np.concatenate([eu_g[out_v], ev_g[out_u]])            # owned endpoint of each cut edge (global id)
        cut_other = np.concatenate([ev_g[out_v], eu_g[out_u]])            # its ghost endpoint
        cut_owner = np.searchsorted(self.bounds, cut_other, side='right') - 1
        self.send_rows = []
        for q in range(world):
            rows = np.unique(cut_owned[cut_owner == q]) - lo if q != rank else np.zeros(0, dtype=np.int64)
            self.send_rows.append(rows.astype(np.int64))
        self.send_counts = np.array([r.size for r in self.send_rows], dtype=np.int64)
        self.send_offsets = np.concatenate([[0], np.cumsum(self.send_counts)])[:-1]
        self.peers = [q for q in range(world) if q != rank and (self.send_counts[q] or self.recv_counts[q])]
        # the same lists per domain (0 nodes, 1 edges, 3 faces): local rows sent to / filled from every rank
        self.hops = 1
        self.send_dom = {0: self.send_rows}
        self.recv_dom = {0: [self.n_own + self.recv_offsets[q] + np.arange(self.recv_counts[q], dtype=np.int64)
                             for q in range(world)]}
        self.faces_local = None

    def set_public_ids(self, i2p):
        """The ranges were cut in a RENUMBERED node order (space-filling curve over the node positions: spatially compact
        parts for graphs whose public numbering has no locality).  i2p[k] = public id of node k of that order; everything
        this object hands out (`global_ids`) or takes (`localise`) stays in PUBLIC ids."""
        if i2p is None:
            return
        self.pub = np.ascontiguousarray(i2p, dtype=np.int64)
        self.p2i = np.empty_like(self.pub)
        self.p2i[self.pub] = np.arange(self.pub.size, dtype=np.int64)

    @property
    def global_ids(self) -> np.ndarray:
        """global (public) node id of every local row (owned, then ghosts)"""
        ids = np.concatenate([np.arange(self.lo, self.hi, dtype=np.int64), self.ghosts])
        return ids if self.pub is None else self.pub[ids]

    def localise(self, global_idx):
        """(mask of entries that are local rows, their local indices) for an array of global (public) node ids"""
        g = np.asarray(global_idx, dtype=np.int64)
        if self.p2i is not None:
            g = self.p2i[g]
        own = (g >= self.lo) & (g < self.hi)
        if self.n_ghost:
            pos = np.minimum(np.searchsorted(self.ghosts, g), self.n_ghost - 1)
            gh = ~own & (self.ghosts[pos] == g)
        else:
            pos, gh = np.zeros(g.shape, dtype=np.int64), np.zeros(g.shape, dtype=bool)
        sel = own | gh
        loc = np.where(own, g - self.lo, self.n_own + pos)
        return sel, loc[sel]

    # ---- constructors ------------------------------------------------------------------------------------
    @staticmethod
    def from_global(rank, world, V, eu, ev, weights=None, align=1, bounds=None) -> 'Partition':
        """generic path: the whole edge list is known on every rank"""
        bounds = split_bounds(V, world, align) if bounds is None else np.asarray(bounds, dtype=np.int64)
        lo, hi = bounds[rank], bounds[rank + 1]
        eu, ev = np.asarray(eu), np.asarray(ev)
        keep = ((eu >= lo) & (eu < hi)) | ((ev >= lo) & (ev < hi))
        w = None if weights is None else np.asarray(weights, dtype=np.float64)[keep]
        return Partition(rank, world, V, bounds, eu[keep], ev[keep], w)

    @staticmethod
    def grid_slab(rank, world, A, B, C=1) -> 'Partition':
        """closed form for row-major grids: node (a, b, c) has id a*B*C + b*C + c and G.edges() yields, per node, its
        +a, +b, +c neighbour (nx.grid_2d_graph(m, n): A=m, B=n, C=1; nx.grid_graph(dim=[d0, d1, d2]): A=d2, B=d1, C=d0;
        SURVEY.md Appendix C).  Only the local slab is generated, so a 10^9-edge grid never exists in one process."""
        plane = B * C
        V = A * plane
        bounds = split_bounds(V, world, align=plane)
        a0, a1 = int(bounds[rank] // plane), int(bounds[rank + 1] // plane)
        lo = a0 * plane
        first = max(a0 - 1, 0) * plane
        ids = np.arange(first, a1 * plane, dtype=np.int64)
        a = ids // plane
        rem = ids - a * plane
        b = rem // C
        c = rem - b * C
        owned = ids >= lo
        valid = np.stack([a + 1 < A, (b + 1 < B) & owned, (c + 1 < C) & owned], axis=1)
        u = np.broadcast_to(ids[:, None], valid.shape)[valid]
        v = (ids[:, None] + np.array([plane, C, 1], dtype=np.int64)[None, :])[valid]
        ghosts = []
        if a0 > 0:
            ghosts.append(np.arange((a0 - 1) * plane, a0 * plane, dtype=np.int64))
        if a1 < A:
            ghosts.append(np.arange(a1 * plane, (a1 + 1) * plane, dtype=np.int64))
        ghosts = np.concatenate(ghosts) if ghosts else np.zeros(0, dtype=np.int64)
        if a1 == a0:
            u, v, ghosts = u[:0], v[:0], ghosts[:0]
        return Partition(rank, world, V, bounds, u, v, None, ghosts=ghosts)


class BlockPartition(Partition):
    """One rank's BLOCK of a row-major grid (2 x 2 x 2 blocks of the 3-D grid instead of slabs -- three faces
    of ~n^2/4 ghost rows per rank instead of two of n^2).  Node (a, b, c) has global id a*B*C + b*C + c and G.edges()
    yields, per node, its +a, +b, +c neighbour (Appendix C), as in `Partition.grid_slab`; only this block and its face
    layers are generated.

    Local rows: [owned rows next to another block (they are sent, and only they read ghost values) | the other owned rows |
    ghost rows], each group in global id order -- so the rows that take part in the exchange form a prefix and the engine
    can run them first and hide the exchange behind the rest (engine.overlap_split).  Every owned row still lists its edges
    in GLOBAL edge order: same summation order, same bits as on one GPU."""

    def __init__(self, rank, blocks, A, B, C=1):
        pa, pb, pc = (int(x) for x in blocks)
        world = pa * pb * pc
        assert 0 <= rank < world and pa <= A and pb <= B and pc <= C
        self.rank, self.world, self.V_global = int(rank), world, A * B * C
        self.blocks, self.dims = (pa, pb, pc), (A, B, C)
        cuts = [np.array([n * k // p for k in range(p + 1)], dtype=np.int64) for n, p in ((A, pa), (B, pb), (C, pc))]
        self.cuts = cuts
        r = (rank // (pb * pc), (rank // pc) % pb, rank % pc)
        lo = [int(cuts[d][r[d]]) for d in range(3)]
        hi = [int(cuts[d][r[d] + 1]) for d in range(3)]
        stride = np.array([B * C, C, 1], dtype=np.int64)
        dims = (A, B, C)

        def box(l, h):
            a, b, c = np.meshgrid(np.arange(l[0], h[0]), np.arange(l[1], h[1]), np.arange(l[2], h[2]), indexing='ij')
            return np.stack([a.ravel(), b.ravel(), c.ravel()], axis=1).astype(np.int64)
        own = box(lo, hi)                                                   # lexicographic = global id order
        own_gid = own @ stride
        # owned rows with a neighbour in another block
        edge_of_block = np.zeros(own.shape[0], dtype=bool)
        ghosts = []
        for d in range(3):
            if lo[d] > 0:
                edge_of_block |= own[:, d] == lo[d]
                l, h = list(lo), list(hi)
                l[d], h[d] = lo[d] - 1, lo[d]
                ghosts.append(box(l, h))
            if hi[d] < dims[d]:
                edge_of_block |= own[:, d] == hi[d] - 1
                l, h = list(lo), list(hi)
                l[d], h[d] = hi[d], hi[d] + 1
                ghosts.append(box(l, h))
        gh = np.concatenate(ghosts) if ghosts else np.zeros((0, 3), dtype=np.int64)
        gh_gid = gh @ stride
        order = np.argsort(gh_gid)
        gh, gh_gid = gh[order], gh_gid[order]
        self.n_own, self.n_ghost = int(own_gid.size), int(gh_gid.size)
        self.n_loc = self.n_own + self.n_ghost
        self.n_boundary = int(edge_of_block.sum())
        self.own_gid = np.concatenate([own_gid[edge_of_block], own_gid[~edge_of_block]])
        self.ghosts = gh_gid
        self.lo, self.hi = 0, 0        # no contiguous global range
        self._sorted_ids = np.concatenate([self.own_gid, self.ghosts])
        self._sort_perm = np.argsort(self._sorted_ids, kind='stable')
        self._sorted_ids = self._sorted_ids[self._sort_perm]
        # edges touching an owned node: out-edges of owned nodes, and the edge a lower ghost sends into the block
        tails = np.concatenate([own, gh])
        tails_owned = np.concatenate([np.ones(self.n_own, dtype=bool), np.zeros(self.n_ghost, dtype=bool)])
        eu, ev, key = [], [], []
        for d in range(3):
            head = tails.copy()
            head[:, d] += 1
            ok = head[:, d] < dims[d]
            inside = np.all((head >= np.array(lo)) & (head < np.array(hi)), axis=1)
            keep = ok & (tails_owned | inside)                            # at least one endpoint is owned
            u, v = tails[keep] @ stride, head[keep] @ stride
            eu.append(u); ev.append(v); key.append(u * 3 + d)               # global edge order: by tail, then direction
        eu, ev, key = np.concatenate(eu), np.concatenate(ev), np.concatenate(key)
        o = np.argsort(key, kind='stable')
        eu, ev = eu[o], ev[o]
        self.edge_u, self.edge_v = self._to_local(eu).astype(np.int32), self._to_local(ev).astype(np.int32)
        self.weights = None
        self.E_loc = int(eu.size)
        # halo lists: ghost rows by owner; owned rows each neighbour needs -- both in global id order
        def owner_of(coords):
            idx = [np.searchsorted(cuts[d], coords[:, d], side='right') - 1 for d in range(3)]
            return idx[0] * (pb * pc) + idx[1] * pc + idx[2]
        gh_owner = owner_of(gh)
        u_own, v_own = self.edge_u < self.n_own, self.edge_v < self.n_own
        cut = u_own != v_own
        cut_owned_local = np.where(u_own[cut], self.edge_u[cut], self.edge_v[cut]).astype(np.int64)
        cut_ghost_local = np.where(u_own[cut], self.edge_v[cut], self.edge_u[cut]).astype(np.int64) - self.n_own
        cut_owner = gh_owner[cut_ghost_local]
        self.send_rows, recv = [], []
        for q in range(world):
            rows = np.unique(cut_owned_local[cut_owner == q]) if q != rank else np.zeros(0, dtype=np.int64)
            rows = rows[np.argsort(self.own_gid[rows], kind='stable')]     # the receiver lists them by global id
            self.send_rows.append(rows.astype(np.int64))
            recv.append((self.n_own + np.nonzero(gh_owner == q)[0]).astype(np.int64) if q != rank else np.zeros(0, dtype=np.int64))
        self.send_counts = np.array([x.size for x in self.send_rows], dtype=np.int64)
        self.recv_counts = np.array([x.size for x in recv], dtype=np.int64)
        self.peers = [q for q in range(world) if q != rank and (self.send_counts[q] or self.recv_counts[q])]
        self.hops = 1
        self.send_dom, self.recv_dom = {0: self.send_rows}, {0: recv}
        self.faces_local = None

    def _to_local(self, gids):
        pos = np.searchsorted(self._sorted_ids, gids)
        assert np.all(self._sorted_ids[np.minimum(pos, self._sorted_ids.size - 1)] == gids), 'node outside this block and its halo'
        return self._sort_perm[pos]

    @property
    def global_ids(self) -> np.ndarray:
        return np.concatenate([self.own_gid, self.ghosts])

    def localise(self, global_idx):
        g = np.asarray(global_idx, dtype=np.int64)
        if self._sorted_ids.size == 0:
            return np.zeros(g.shape, dtype=bool), np.zeros(0, dtype=np.int64)
        pos = np.minimum(np.searchsorted(self._sorted_ids, g), self._sorted_ids.size - 1)
        sel = self._sorted_ids[pos] == g
        return sel, self._sort_perm[pos][sel]


def _reach(V, eu, ev, lo, hi, hops):
    """node levels: 0 for [lo, hi), k for nodes first reached after k hops, -1 beyond `hops`"""
    level = np.full(V, -1, dtype=np.int64)
    level[lo:hi] = 0
    for k in range(1, hops + 1):
        inside = level >= 0
        m = inside[eu] != inside[ev]
        new = np.concatenate([eu[m & ~inside[eu]], ev[m & ~inside[ev]]])
        level[new[level[new] < 0]] = k
    return level


class DeepPartition(Partition):
    """One rank's part of a graph with `hops` layers of ghost nodes and every edge / face that lies inside them, for
    edge and face fields and for laws that chain several gathers: the intermediate vectors (div, curl, upwind
    aggregates) are recomputed on the ghost layers instead of exchanged, so one exchange of the STATE per stage is
    enough.  Nodes are owned by index range, an edge by the owner of its tail node, a face by the owner
    of its first node.  Local edges and faces keep their global relative order (rows sum in the same order as on one
    GPU); local nodes are [owned | ghosts by global id]."""

    def __init__(self, rank, world, V, eu, ev, weights, bounds, hops, face_ptr=None, face_nodes=None):
        self.rank, self.world, self.V_global = int(rank), int(world), int(V)
        self.bounds = np.asarray(bounds, dtype=np.int64)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi, self.n_own, self.hops = lo, hi, hi - lo, int(hops)
        eu, ev = np.asarray(eu, dtype=np.int64), np.asarray(ev, dtype=np.int64)
        F = 0 if face_ptr is None else len(face_ptr) - 1
        fp = np.zeros(1, dtype=np.int64) if face_ptr is None else np.asarray(face_ptr, dtype=np.int64)
        fn = np.zeros(0, dtype=np.int64) if face_nodes is None else np.asarray(face_nodes, dtype=np.int64)
        node_owner = np.searchsorted(self.bounds, np.arange(V), side='right') - 1
        edge_owner = node_owner[eu]
        face_owner = node_owner[fn[fp[:-1]]] if F else np.zeros(0, dtype=np.int64)

        def local_sets(r):
            lev = _reach(V, eu, ev, int(self.bounds[r]), int(self.bounds[r + 1]), self.hops)
            inn = lev >= 0
            e_in = inn[eu] & inn[ev]
            f_in = np.minimum.reduceat(inn[fn].astype(np.int8), fp[:-1]).astype(bool) if F else np.zeros(0, dtype=bool)
            return lev, inn, e_in, f_in
        lev, inn, e_in, f_in = local_sets(rank)
        ghosts = np.nonzero(inn & (node_owner != rank))[0]
        self.ghosts, self.n_ghost = ghosts, int(ghosts.size)
        self.n_loc = self.n_own + self.n_ghost
        g2l = np.full(V, -1, dtype=np.int64)
        g2l[lo:hi] = np.arange(self.n_own)
        g2l[ghosts] = self.n_own + np.arange(self.n_ghost)
        self.edge_gid = np.nonzero(e_in)[0]
        self.face_gid = np.nonzero(f_in)[0]
        self.edge_u, self.edge_v = g2l[eu[self.edge_gid]].astype(np.int32), g2l[ev[self.edge_gid]].astype(np.int32)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)[self.edge_gid]
        self.E_loc, self.F_loc = int(self.edge_gid.size), int(self.face_gid.size)
        if F:
            sizes = (fp[1:] - fp[:-1])[self.face_gid]
            self.face_ptr_l = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            # positions of the kept faces' nodes in fn: each face's start repeated, plus the running offset inside it
            take = (np.repeat(fp[self.face_gid] - self.face_ptr_l[:-1], sizes) + np.arange(int(self.face_ptr_l[-1]))
                    if self.F_loc else np.zeros(0, dtype=np.int64))
            self.face_nodes_l = g2l[fn[take]].astype(np.int32)
            self.max_face = int((fp[1:] - fp[:-1]).max())
        else:
            self.face_ptr_l, self.face_nodes_l, self.max_face = np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), 0
        self.level_node = lev[self.global_ids]
        # rows filled from / sent to every rank, per domain, in global order inside an owner
        gid = {0: self.global_ids, 1: self.edge_gid, 3: self.face_gid}
        own = {0: node_owner[gid[0]], 1: edge_owner[gid[1]], 3: face_owner[gid[3]] if F else np.zeros(0, dtype=np.int64)}
        self.gid_dom = gid
        self.owned_dom = {d: own[d] == rank for d in gid}
        self.recv_dom = {d: [np.nonzero(own[d] == q)[0].astype(np.int64) if q != rank else np.zeros(0, dtype=np.int64)
                             for q in range(world)] for d in gid}
        self.send_dom = {d: [np.zeros(0, dtype=np.int64) for _ in range(world)] for d in gid}
        e_g2l = np.full(eu.size, -1, dtype=np.int64); e_g2l[self.edge_gid] = np.arange(self.E_loc)
        f_g2l = np.full(F, -1, dtype=np.int64); f_g2l[self.face_gid] = np.arange(self.F_loc)
        # Reach is symmetric (undirected graph, same number of hops): a row of mine can only lie inside rank q's layers if
        # a node of q lies inside mine -- so the only ranks that receive from this one are the owners of its ghost nodes,
        # not all `world - 1` others (their layers are not walked at all).
        for q in np.unique(node_owner[ghosts]).tolist():
            _, inn_q, e_q, f_q = local_sets(q)
            self.send_dom[0][q] = g2l[np.nonzero(inn_q & (node_owner == rank))[0]]
            self.send_dom[1][q] = e_g2l[np.nonzero(e_q & (edge_owner == rank))[0]]
            if F:
                self.send_dom[3][q] = f_g2l[np.nonzero(f_q & (face_owner == rank))[0]]
        for d in gid:
            for q in range(world):
                assert np.all(self.send_dom[d][q] >= 0), 'a rank must hold every row it owns'
        self.send_rows = self.send_dom[0]
        self.peers = [q for q in range(world) if q != rank and
                      any(self.send_dom[d][q].size or self.recv_dom[d][q].size for d in gid)]
        self.faces_local = (self.face_ptr_l, self.face_nodes_l)

    def localise(self, global_idx, domain=0):
        g = np.asarray(global_idx, dtype=np.int64)
        ids = self.gid_dom[domain]
        order = np.argsort(ids, kind='stable')
        pos = np.minimum(np.searchsorted(ids[order], g), max(ids.size - 1, 0))
        sel = (ids[order][pos] == g) if ids.size else np.zeros(g.shape, dtype=bool)
        return sel, order[pos][sel]


class HaloExchange:
    """per-stage exchange of boundary values: point-to-point over torch.distributed (NCCL on GPUs -> NVLink;
    gloo in the CPU tests).  `blocks` lists the state blocks that travel, as (offset, domain) pairs (a bare offset is a
    node block); `send` / `recv` are contiguous tensors laid out by peer and, inside a peer, by block."""

    def __init__(self, part: Partition, blocks=None, nblocks=None, group=None):
        if blocks is None:
            blocks = [0] * int(nblocks if nblocks is not None else 1)
        self.part, self.group = part, group
        self.blocks = [(b, 0) if np.isscalar(b) else (int(b[0]), int(b[1])) for b in blocks]
        self.nblocks = len(self.blocks)
        w = part.world
        self.send_sizes = np.array([sum(part.send_dom[d][q].size for _, d in self.blocks) for q in range(w)], dtype=np.int64)
        self.recv_sizes = np.array([sum(part.recv_dom[d][q].size for _, d in self.blocks) for q in range(w)], dtype=np.int64)
        self.send_off = np.concatenate([[0], np.cumsum(self.send_sizes)])[:-1]
        self.recv_off = np.concatenate([[0], np.cumsum(self.recv_sizes)])[:-1]
        self.n_send, self.n_recv = int(self.send_sizes.sum()), int(self.recv_sizes.sum())

    def state_indices(self, block_offsets, which):
        """positions in the state vector of the rows sent ('send') or of the ghost rows filled ('recv'), grouped by
        peer and, inside a peer, by state block (`block_offsets`: offsets, or (offset, domain) pairs)"""
        p, out = self.part, []
        blocks = [(b, 0) if np.isscalar(b) else (int(b[0]), int(b[1])) for b in block_offsets]
        lists = p.send_dom if which == 'send' else p.recv_dom
        for q in range(p.world):
            for off, d in blocks:
                out.append(off + lists[d][q])
        return np.concatenate(out).astype(np.int64) if out else np.zeros(0, dtype=np.int64)

    def exchange(self, send, recv):
        """enqueue the transfers; returns the request list (wait() orders later work on the current stream)"""
        import torch.distributed as dist
        if send.is_cuda and dist.get_backend(self.group) == 'gloo':
            # test transport (several ranks sharing one GPU, where NCCL cannot run): stage through the host, blocking
            import torch
            send_h, recv_h = send.cpu(), torch.empty(recv.shape, dtype=recv.dtype)
            for r in self.exchange(send_h, recv_h):
                r.wait()
            recv.copy_(recv_h)
            return []
        p, ops = self.part, []
        for q in p.peers:
            so, sc = int(self.send_off[q]), int(self.send_sizes[q])
            ro, rc = int(self.recv_off[q]), int(self.recv_sizes[q])
            if rc:
                ops.append(dist.P2POp(dist.irecv, recv[ro:ro + rc], q, self.group))
            if sc:
                ops.append(dist.P2POp(dist.isend, send[so:so + sc], q, self.group))
        return dist.batch_isend_irecv(ops) if ops else []
