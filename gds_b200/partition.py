"""Graph partitioning for multi-GPU stepping -- no counterpart in the single-process reference.

Nodes are split into contiguous index ranges (slabs of a row-major lattice; `align` keeps cuts on plane boundaries).
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

    @property
    def global_ids(self) -> np.ndarray:
        """global node id of every local row (owned, then ghosts)"""
        return np.concatenate([np.arange(self.lo, self.hi, dtype=np.int64), self.ghosts])

    def localise(self, global_idx):
        """(mask of entries that are local rows, their local indices) for an array of global node ids"""
        g = np.asarray(global_idx, dtype=np.int64)
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


class HaloExchange:
    """per-stage exchange of boundary values: point-to-point over torch.distributed (NCCL on GPUs -> NVLink;
    gloo in the CPU tests).  `send` / `recv` are contiguous tensors laid out by peer (Partition.send_offsets /
    recv_offsets), one segment of `nblocks` values per listed row."""

    def __init__(self, part: Partition, nblocks: int = 1, group=None):
        self.part, self.nblocks, self.group = part, int(nblocks), group
        self.n_send = int(part.send_counts.sum()) * self.nblocks
        self.n_recv = int(part.recv_counts.sum()) * self.nblocks

    def state_indices(self, block_offsets, which):
        """positions in the state vector of the rows sent ('send') or of the ghost rows filled ('recv'), grouped by
        peer and, inside a peer, by state block"""
        p, out = self.part, []
        for q in range(p.world):
            if which == 'send':
                rows = p.send_rows[q]
            else:
                rows = p.n_own + p.recv_offsets[q] + np.arange(p.recv_counts[q], dtype=np.int64)
            for off in block_offsets:
                out.append(off + rows)
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
        p, nb, ops = self.part, self.nblocks, []
        for q in p.peers:
            so, sc = int(p.send_offsets[q]) * nb, int(p.send_counts[q]) * nb
            ro, rc = int(p.recv_offsets[q]) * nb, int(p.recv_counts[q]) * nb
            if rc:
                ops.append(dist.P2POp(dist.irecv, recv[ro:ro + rc], q, self.group))
            if sc:
                ops.append(dist.P2POp(dist.isend, send[so:so + sc], q, self.group))
        return dist.batch_isend_irecv(ops) if ops else []
