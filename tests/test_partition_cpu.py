"""Partitioning for multi-GPU stepping (gds_b200/partition.py): closed-form slabs against the generic path, halo-list
symmetry, partition-independent summation order of the lowered rows, and a 2-rank halo exchange over gloo."""
import os

import networkx as nx
import numpy as np
import pytest

import gds_b200
from gds_b200.graph import NativeGraph
from gds_b200.partition import HaloExchange, Partition, split_bounds


def global_edges(kind, *dims):
    hg = NativeGraph(kind, *dims).hostgraph()
    eu, ev = hg.edges()
    return hg.V, eu.astype(np.int64), ev.astype(np.int64)


@pytest.mark.parametrize('kind,dims,abc', [('grid_2d', (9, 5), (9, 5, 1)), ('grid_2d', (4, 7), (4, 7, 1)),
                                           ('grid', (3, 4, 6), (6, 4, 3)), ('grid', (5, 2, 8), (8, 2, 5))])
@pytest.mark.parametrize('world', [1, 2, 3, 4])
def test_grid_slab_equals_generic_partition(kind, dims, abc, world):
    V, eu, ev = global_edges(kind, *dims)
    A, B, C = abc
    for rank in range(world):
        a = Partition.grid_slab(rank, world, A, B, C)
        b = Partition.from_global(rank, world, V, eu, ev, None, bounds=split_bounds(V, world, align=B * C))
        assert (a.lo, a.hi) == (b.lo, b.hi)
        assert np.array_equal(a.ghosts, b.ghosts)
        assert np.array_equal(a.edge_u, b.edge_u) and np.array_equal(a.edge_v, b.edge_v)
        for q in range(world):
            assert np.array_equal(a.send_rows[q], b.send_rows[q])
        assert np.array_equal(a.recv_counts, b.recv_counts)


@pytest.mark.parametrize('world', [2, 3, 5])
def test_halo_lists_are_symmetric(world):
    G = nx.random_geometric_graph(300, 0.12, seed=5)
    V = G.number_of_nodes()
    eu = np.array([e[0] for e in G.edges()], dtype=np.int64)
    ev = np.array([e[1] for e in G.edges()], dtype=np.int64)
    parts = [Partition.from_global(r, world, V, eu, ev) for r in range(world)]
    assert sum(p.n_own for p in parts) == V
    for r, p in enumerate(parts):
        for q in range(world):
            if q == r:
                continue
            # what r sends to q (global ids) is exactly the part of q's ghost segment owned by r, in the same order
            sent = p.send_rows[q] + p.lo
            o, c = parts[q].recv_offsets[r], parts[q].recv_counts[r]
            assert np.array_equal(sent, parts[q].ghosts[o:o + c])
        # every local edge has an owned endpoint; every cut edge appears on both sides
        assert np.all((p.edge_u < p.n_own) | (p.edge_v < p.n_own))
    total_local = sum(p.E_loc for p in parts)
    cut = sum(int(np.sum((p.edge_u >= p.n_own) | (p.edge_v >= p.n_own))) for p in parts) // 2
    assert total_local == eu.size + cut


def test_owned_rows_keep_the_global_summation_order(monkeypatch):
    """the lowered B1 row of an owned node lists the same neighbours in the same order as the undistributed graph"""
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    try:
        G = nx.grid_2d_graph(7, 6)
        full = gds_b200.LoweredGraph.of(nx.grid_2d_graph(7, 6))
        p0, i0, s0 = full.dev.export_csr(0)
        eu0, ev0 = full.edge_u, full.edge_v
        for rank in range(3):
            Gd = gds_b200.distribute(nx.grid_2d_graph(7, 6), rank, 3, align=6)
            low = gds_b200.LoweredGraph.of(Gd)
            part = low.part
            gid = part.global_ids
            p, i, s = low.dev.export_csr(0)
            for r in range(part.n_own):
                g = part.lo + r
                other_g = [eu0[e] if ev0[e] == g else ev0[e] for e in i0[p0[g]:p0[g + 1]]]
                other_l = [gid[low.edge_u[e]] if low.edge_v[e] == r else gid[low.edge_v[e]] for e in i[p[r]:p[r + 1]]]
                assert other_g == other_l
                assert np.array_equal(s0[p0[g]:p0[g + 1]], s[p[r]:p[r + 1]])
            # node labels of the local rows
            assert list(low.nodes.keys()) == [list(G.nodes())[int(k)] for k in gid]
    finally:
        gds_b200.Context._default.clear()


def test_boundary_spec_localisation():
    part = Partition.grid_slab(1, 3, 9, 4, 1)       # owns rows 3..5 of a 9 x 4 grid: ids 12..23, ghosts 8..11 and 24..27
    spec = gds_b200.BoundarySpec(np.array([0, 9, 13, 22, 26, 35]), fun=lambda t: t + np.arange(6.0))
    loc = spec.localised(part)
    assert np.array_equal(loc.indices, [12 + 1, 1, 10, 12 + 4 + 2])      # ghost 9, owned 13 and 22, ghost 26
    assert np.array_equal(loc.fun(1.0), [2.0, 3.0, 4.0, 5.0])


# ------------------------------------------------------------------------------------------------
# two ranks over gloo: halo exchange + partitioned explicit stepping against the undistributed result
# ------------------------------------------------------------------------------------------------
def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        A, B, C = 6, 5, 4
        part = Partition.grid_slab(rank, world, A, B, C)
        V = A * B * C
        rng = np.random.default_rng(0)
        x_global = rng.standard_normal(V)
        hx = HaloExchange(part, nblocks=1)
        send_idx, recv_idx = hx.state_indices([0], 'send'), hx.state_indices([0], 'recv')
        x = np.zeros(part.n_loc)
        x[:part.n_own] = x_global[part.lo:part.hi]
        lap = lambda y: (np.bincount(part.edge_u, weights=y[part.edge_v] - y[part.edge_u], minlength=part.n_loc) +
                         np.bincount(part.edge_v, weights=y[part.edge_u] - y[part.edge_v], minlength=part.n_loc))
        for step in range(4):
            send = torch.from_numpy(x[send_idx].copy())
            recv = torch.empty(hx.n_recv, dtype=torch.float64)
            for r in hx.exchange(send, recv):
                r.wait()
            x[recv_idx] = recv.numpy()
            if step == 0:
                assert np.array_equal(x, x_global[part.global_ids])
            x[:part.n_own] = (x + 0.05 * lap(x))[:part.n_own]
        ret[rank] = (part.lo, part.hi, x[:part.n_own].copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_halo_exchange_gloo():
    import torch.multiprocessing as mp
    world, port = 2, 29000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    # undistributed reference of the same 3 explicit steps
    A, B, C = 6, 5, 4
    V, eu, ev = global_edges('grid', C, B, A)
    x = np.random.default_rng(0).standard_normal(V)
    for _ in range(4):
        d = (np.bincount(eu, weights=x[ev] - x[eu], minlength=V) + np.bincount(ev, weights=x[eu] - x[ev], minlength=V))
        x = x + 0.05 * d
    got = np.empty(V)
    for r in range(world):
        lo, hi, vals = ret[r]
        got[lo:hi] = vals
    assert np.allclose(got, x, rtol=0, atol=1e-14)


def test_peer_push_lists_fill_every_ghost_row():
    """the peer-to-peer halo stores state[src_idx] of rank r into state[dst_idx] of rank q, where dst_idx is what q
    publishes as the positions of its ghost rows owned by r (engine._attach_peers): emulate the push with numpy for a
    two-block state (order 2) and check that every ghost row ends up with its owner's value"""
    world, V = 4, 9 * 6 * 5
    parts = [Partition.grid_slab(r, world, 9, 6, 5) for r in range(world)]
    rng = np.random.default_rng(3)
    xg = [rng.standard_normal(V), rng.standard_normal(V)]                     # two state blocks, global order
    state, blocks = [], []
    for p in parts:
        blocks.append([0, p.n_loc])                                            # block offsets in the local state
        st = np.full(2 * p.n_loc, np.nan)
        for b in range(2):
            st[blocks[-1][b]:blocks[-1][b] + p.n_own] = xg[b][p.lo:p.hi]       # owned rows only
        state.append(st)
    published = []
    for r, p in enumerate(parts):
        published.append({q: np.concatenate([off + p.n_own + p.recv_offsets[q] + np.arange(p.recv_counts[q])
                                             for off in blocks[r]]) for q in p.peers})
    for r, p in enumerate(parts):
        for q in p.peers:
            src = np.concatenate([off + p.send_rows[q] for off in blocks[r]])
            dst = published[q][r]
            assert src.size == dst.size
            state[q][dst] = state[r][src]                                       # the push kernel
    for r, p in enumerate(parts):
        gid = p.global_ids
        for b in range(2):
            assert np.array_equal(state[r][blocks[r][b]:blocks[r][b] + p.n_loc], xg[b][gid])


def _hodge1(V, eu, ev, faces, y):
    """-B1^T B1 y - B2^T B2 y with unit weights, faces as lists of node ids (the sign rule of gds.py:227-236)"""
    import scipy.sparse as sp
    E = eu.size
    ar = np.arange(E)
    B1 = sp.csr_matrix((np.concatenate([-np.ones(E), np.ones(E)]), (np.concatenate([eu, ev]), np.concatenate([ar, ar]))),
                       shape=(V, E))
    eid = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(eu, ev))}
    rows, cols, vals = [], [], []
    for fi, f in enumerate(faces):
        k = len(f)
        for i in range(k):
            a, b = int(f[i]), int(f[(i + 1) % k])
            if (a, b) in eid:
                rows.append(fi); cols.append(eid[(a, b)]); vals.append(1.0)
            elif (b, a) in eid:
                rows.append(fi); cols.append(eid[(b, a)]); vals.append(-1.0)
    B2 = sp.csr_matrix((vals, (rows, cols)), shape=(max(len(faces), 1), E))
    return -(B1.T @ (B1 @ y)) - (B2.T @ (B2 @ y))


@pytest.mark.parametrize('kind,dims,hops,world', [('hexagonal', (5, 6), 4, 2), ('hexagonal', (5, 6), 4, 3), ('triangular', (6, 8), 2, 2),
                                                  ('triangular', (6, 8), 2, 3),
                                                  ('hexagonal', (6, 30), 3, 7), ('triangular', (8, 40), 2, 6)])   # most ranks are not neighbours
def test_deep_partition_recomputes_two_hop_operators_on_the_ghost_layers(kind, dims, hops, world):
    """edge fields on a partitioned graph: with enough ghost layers the Hodge-1 Laplacian of the OWNED edges needs no
    exchange of intermediates -- every rank evaluates it on its local graph after one exchange of the edge values"""
    from gds_b200.partition import DeepPartition
    hg = NativeGraph(kind, *dims).hostgraph(with_faces=True)
    eu, ev = (a.astype(np.int64) for a in hg.edges())
    fp, fn = hg.faces()
    faces = [fn[fp[i]:fp[i + 1]] for i in range(fp.size - 1)]
    V, E = hg.V, eu.size
    y = np.random.default_rng(4).standard_normal(E)
    want = _hodge1(V, eu, ev, faces, y)
    bounds = split_bounds(V, world)
    parts = [DeepPartition(r, world, V, eu, ev, None, bounds, hops, fp, fn) for r in range(world)]
    if world > 3:
        assert any(len(p.peers) < world - 1 for p in parts)
    # ownership partitions every domain, halo lists are symmetric
    for d, n in ((0, V), (1, E), (3, len(faces))):
        owned = np.concatenate([p.gid_dom[d][p.owned_dom[d]] for p in parts])
        assert np.array_equal(np.sort(owned), np.arange(n))
        for r, p in enumerate(parts):
            for q in range(world):
                if q != r:
                    assert np.array_equal(p.gid_dom[d][p.send_dom[d][q]], parts[q].gid_dom[d][parts[q].recv_dom[d][r]])
    # exchange the edge values, evaluate locally, compare the owned rows
    local = []
    for p in parts:
        v = np.full(p.E_loc, np.nan)
        v[p.owned_dom[1]] = y[p.edge_gid[p.owned_dom[1]]]
        local.append(v)
    for r, p in enumerate(parts):
        for q in p.peers:
            local[q][parts[q].recv_dom[1][r]] = local[r][p.send_dom[1][q]]
    for p, v in zip(parts, local):
        assert np.array_equal(v, y[p.edge_gid])
        lf = [p.face_nodes_l[p.face_ptr_l[i]:p.face_ptr_l[i + 1]] for i in range(p.F_loc)]
        got = _hodge1(p.n_loc, p.edge_u.astype(np.int64), p.edge_v.astype(np.int64), lf, v)
        own = p.owned_dom[1]
        assert np.allclose(got[own], want[p.edge_gid[own]], rtol=0, atol=1e-12)


@pytest.mark.parametrize('world,hops', [(2, 2), (3, 2), (4, 3)])
def test_deep_partition_is_enough_for_the_edge_self_advection(world, hops):
    """The two-pass nonlinear operator (node aggregates with weighted degrees, then the edge formula; gds.py:326-345) on
    an IRREGULAR weighted graph: with two ghost layers every rank gets its OWNED edges right from its local graph alone
    -- the aggregates at the far endpoint of an owned edge need that endpoint's whole star, one layer further out."""
    import networkx as nx
    from oracle import gds_oracle
    from gds_b200.partition import DeepPartition
    G = nx.random_geometric_graph(260, 0.11, seed=9)
    G = nx.convert_node_labels_to_integers(G.subgraph(max(nx.connected_components(G), key=len)).copy())
    # number the nodes along x so that index ranges are spatial slabs
    order = sorted(G.nodes(), key=lambda n: G.nodes[n]['pos'][0])
    G = nx.relabel_nodes(G, {n: i for i, n in enumerate(order)})
    H = nx.Graph()
    H.add_nodes_from(range(G.number_of_nodes()))
    rng = np.random.default_rng(2)
    H.add_weighted_edges_from((min(u, v), max(u, v), float(rng.uniform(0.5, 2.0))) for u, v in G.edges())
    H.faces = []
    full = gds_oracle.edge_gds(H)
    eu = np.array([u for u, v in H.edges()], dtype=np.int64)
    ev = np.array([v for u, v in H.edges()], dtype=np.int64)
    w = np.array([d['weight'] for _, _, d in H.edges(data=True)])
    V, E = H.number_of_nodes(), eu.size
    y = rng.standard_normal(E)
    y[::9] = 0.0
    want = full.advect(y)
    bounds = split_bounds(V, world)
    def local_advect(p):
        """the oracle's operator on the rank's local graph, in the partition's own edge order and orientation
        (networkx lists the edges of a rebuilt graph in adjacency order, possibly reversed: an edge vector changes
        sign with the orientation of its edge)"""
        L = nx.Graph()
        L.add_nodes_from(range(p.n_loc))
        L.add_weighted_edges_from((int(a), int(b), float(c)) for a, b, c in zip(p.edge_u, p.edge_v, p.weights))
        L.faces = []
        loc = gds_oracle.edge_gds(L)
        keys = list(loc.edges.keys())
        idx = np.array([loc.edges[(int(a), int(b))] for a, b in zip(p.edge_u, p.edge_v)])
        sgn = np.array([1.0 if keys[j] == (int(a), int(b)) else -1.0 for j, a, b in zip(idx, p.edge_u, p.edge_v)])
        y_loc = np.zeros(p.E_loc)
        y_loc[idx] = sgn * y[p.edge_gid]
        return sgn * loc.advect(y_loc)[idx]

    one_layer_ok = []
    for r in range(world):
        p = DeepPartition(r, world, V, eu, ev, w, bounds, hops)
        got, own = local_advect(p), p.owned_dom[1]
        assert own.any() and np.allclose(got[own], want[p.edge_gid[own]], rtol=0, atol=1e-12)
        p1 = DeepPartition(r, world, V, eu, ev, w, bounds, 1)
        o1 = p1.owned_dom[1]
        one_layer_ok.append(np.allclose(local_advect(p1)[o1], want[p1.edge_gid[o1]], rtol=0, atol=1e-9))
    # one layer is NOT enough for a rank that owns a cut edge (its far endpoint's star is incomplete); the last rank owns
    # none (an edge belongs to the owner of its tail = lower index) and gets away with it
    assert not all(one_layer_ok) and one_layer_ok[-1]


@pytest.mark.parametrize('world', [2, 4, 7])
def test_graphs_without_index_locality_are_cut_along_a_space_filling_curve(world, monkeypatch):
    """random geometric graph: node ids are random in space, so index ranges of the PUBLIC numbering would cut almost every
    edge; the ranges are taken in a space-filling-curve order instead (spatially compact parts), while every id handed
    out or taken by the partition stays public"""
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    try:
        n, r = 4000, 0.04
        hg = NativeGraph('random_geometric', n, r, 42).hostgraph()
        eu, ev = [a.astype(np.int64) for a in hg.edges()]
        owner = np.full(n, -1)
        cut_curve = 0
        for rank in range(world):
            G = gds_b200.distribute(NativeGraph('random_geometric', n, r, 42), rank, world, align=32)
            low = gds_b200.LoweredGraph.of(G)
            part = low.part
            gid = part.global_ids
            assert part.pub is not None and len(set(gid.tolist())) == gid.size
            owner[gid[:part.n_own]] = rank
            # local edges = every global edge touching an owned node, in global edge order, public endpoints preserved
            own = np.zeros(n, dtype=bool); own[gid[:part.n_own]] = True
            keep = own[eu] | own[ev]
            assert np.array_equal(gid[low.edge_u], eu[keep]) and np.array_equal(gid[low.edge_v], ev[keep])
            cut_curve += int(np.sum(own[eu] != own[ev]))
            sel, loc = part.localise(np.arange(n))
            assert np.array_equal(gid[loc], np.arange(n)[sel]) and sel.sum() == part.n_loc
        assert np.all(owner >= 0)
        cut_curve //= 2
        cut_index = int(np.sum((eu * world // n) != (ev * world // n)))      # index ranges of the public numbering
        assert cut_curve < 0.25 * cut_index and cut_curve < 0.2 * eu.size, (cut_curve, cut_index, eu.size)
    finally:
        gds_b200.Context._default.clear()


@pytest.mark.parametrize('world', [2, 5])
def test_graphs_without_positions_are_cut_along_breadth_first_levels(world, monkeypatch):
    """a lattice whose nodes were relabelled at random and that carries no positions: index ranges of the public numbering
    cut almost every edge; the ranges are taken in reverse Cuthill-McKee order instead, ids stay public"""
    import networkx as nx
    monkeypatch.setenv('GDS_B200_DEVICE', '-1')
    gds_b200.Context._default.clear()
    try:
        H = nx.grid_2d_graph(40, 50)
        perm = np.random.default_rng(6).permutation(H.number_of_nodes())
        G0 = nx.relabel_nodes(H, {v: int(perm[i]) for i, v in enumerate(H.nodes())})
        def make():
            G = nx.Graph()
            G.add_nodes_from(sorted(G0.nodes()))          # public id = label: no locality at all
            G.add_edges_from(G0.edges())
            G.faces = []
            return G
        n = H.number_of_nodes()
        ref = make()
        eu = np.array([a for a, b in ref.edges()]); ev = np.array([b for a, b in ref.edges()])
        owner, cut = np.full(n, -1), 0
        for rank in range(world):
            low = gds_b200.LoweredGraph.of(gds_b200.distribute(make(), rank, world))
            part = low.part
            gid = part.global_ids
            assert part.pub is not None and len(set(gid.tolist())) == gid.size
            owner[gid[:part.n_own]] = rank
            own = np.zeros(n, dtype=bool); own[gid[:part.n_own]] = True
            keep = own[eu] | own[ev]
            assert np.array_equal(gid[low.edge_u], eu[keep]) and np.array_equal(gid[low.edge_v], ev[keep])
            cut += int(np.sum(own[eu] != own[ev]))
        assert np.all(owner >= 0)
        cut //= 2
        cut_index = int(np.sum((eu * world // n) != (ev * world // n)))
        assert cut < 0.2 * cut_index and cut < 0.12 * eu.size, (cut, cut_index, eu.size)
    finally:
        gds_b200.Context._default.clear()


def test_overlap_split_covers_every_boundary_row():
    """engine.overlap_split: the two row ranges [0, lo) and [hi, n_own) hold every row adjacent to a cut, are multiples of
    128, as short as possible, and the split is refused when there is nothing to hide an exchange behind"""
    from gds_b200.engine import overlap_split
    n = 7072
    both = np.concatenate([np.arange(n), np.arange(n * n - n, n * n)])
    assert overlap_split(both, n * n) == (7168, n * n - 7168)            # one lattice row per side, rounded to 128
    assert overlap_split(np.arange(n), n * n) == (7168, n * n)             # first rank of a chain: nothing at the top
    assert overlap_split(np.arange(n * n - n, n * n), n * n) == (0, n * n - 7168)
    assert overlap_split(np.arange(0, n * n, 50), n * n) is None           # boundary rows everywhere (unordered graph)
    assert overlap_split(np.zeros(0, dtype=np.int64), n * n) is None
    assert overlap_split(np.arange(10), 500) is None                        # too small to bother
    rng = np.random.default_rng(0)
    for _ in range(20):
        n_own = int(rng.integers(2000, 100000))
        k = int(rng.integers(1, 40))
        rows = np.unique(np.concatenate([rng.integers(0, 300, k), rng.integers(n_own - 300, n_own, k)]))
        lo, hi = overlap_split(rows, n_own)
        assert lo % 128 == 0 and hi % 128 == 0 and 0 <= lo <= hi <= n_own
        assert np.all((rows < lo) | (rows >= hi))


def test_stage_independent_laws_are_recognised():
    """engine._reads_stage: a law is collapsible when its expression contains neither the stage vector nor the stage time"""
    from gds_b200.engine import _reads_stage
    from gds_b200.trace import Gather, Stage, State, Time, as_expr

    class F:   # stand-in for a field
        _domain_code, _low = 0, None
    f = F()
    lap_acc = Gather.__new__(Gather)
    lap_acc.args, lap_acc.domain, lap_acc.graph = (State(f),), 0, None
    lap_stage = Gather.__new__(Gather)
    lap_stage.args, lap_stage.domain, lap_stage.graph = (Stage(f),), 0, None
    assert not _reads_stage(2.0 * lap_acc + State(f))
    assert _reads_stage(2.0 * lap_stage)
    assert _reads_stage(lap_acc * Time())
    assert _reads_stage(lap_acc + 0.1 * Stage(f))


@pytest.mark.parametrize('kind,dims,abc,blocks', [('grid', (5, 6, 7), (7, 6, 5), (2, 2, 2)), ('grid', (4, 9, 6), (6, 9, 4), (3, 2, 1)),
                                                   ('grid_2d', (9, 8), (9, 8, 1), (2, 3, 1)), ('grid', (3, 3, 8), (8, 3, 3), (4, 1, 1))])
def test_block_partition_of_a_grid(kind, dims, abc, blocks):
    """BlockPartition: every node owned once; local edges = every global edge touching an owned node, in global edge order
    with public endpoints preserved; halo lists symmetric and in the same order on both sides; the rows that take part
    in the exchange form a prefix of the local numbering"""
    from gds_b200.partition import BlockPartition
    V, eu, ev = global_edges(kind, *dims)
    A, B, C = abc
    world = int(np.prod(blocks))
    parts = [BlockPartition(r, blocks, A, B, C) for r in range(world)]
    owner = np.full(V, -1)
    for r, p in enumerate(parts):
        gid = p.global_ids
        assert np.all(owner[gid[:p.n_own]] == -1)
        owner[gid[:p.n_own]] = r
    assert np.all(owner >= 0)
    for r, p in enumerate(parts):
        gid = p.global_ids
        own = owner == r
        keep = own[eu] | own[ev]
        assert np.array_equal(gid[p.edge_u], eu[keep]) and np.array_equal(gid[p.edge_v], ev[keep])
        assert np.all(owner[gid[p.n_own:]] != r)
        # rows with a cut edge are exactly the first n_boundary rows
        cut_rows = np.unique(np.concatenate([p.edge_u[(p.edge_v >= p.n_own)], p.edge_v[(p.edge_u >= p.n_own)]]))
        assert np.array_equal(cut_rows, np.arange(p.n_boundary))
        for q in p.peers:
            sent = gid[p.send_rows[q]]
            got = parts[q].global_ids[parts[q].recv_dom[0][r]]
            assert np.array_equal(sent, got) and sent.size > 0
        sel, loc = p.localise(np.arange(V))
        assert np.array_equal(gid[loc], np.arange(V)[sel]) and sel.sum() == p.n_loc
        assert len(p.peers) <= 6
