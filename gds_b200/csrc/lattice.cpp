// Native graph construction for sizes where an nx.Graph is impossible (10^8 .. 10^9 edges):
//   * lattice generators that reproduce networkx's node order, edge order/orientation (G.edges()) and adjacency
//     insertion order exactly (networkx/generators/lattice.py: grid_2d_graph, grid_graph, triangular_lattice_graph,
//     hexagonal_lattice_graph; networkx/classes/reportviews.py EdgeView iteration) -- SURVEY.md Appendix C;
//   * random_geometric_graph (networkx/generators/geometric.py) with Python's Mersenne-Twister stream;
//   * the planar face walk of gds/utils/graph/generators.py:444-511.
// Host-only code (no CUDA).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "internal.h"

struct gds_hostgraph {
  int64_t V = 0, E = 0, F = 0;
  int label_dim = 0;                 // 2 or 3 ints per node label (0: labels are 0..V-1)
  std::vector<int32_t> labels;       // [V x label_dim]
  std::vector<double> pos;           // [V x 2] or empty
  std::vector<int32_t> eu, ev;       // G.edges() order, stored orientation
  std::vector<int64_t> adj_ptr;      // adjacency in G.neighbors() order
  std::vector<int32_t> adj_nbr;
  std::vector<int64_t> face_ptr;     // faces (outer face removed) in discovery order
  std::vector<int32_t> face_nodes;
};

// ---------------------------------------------------------------------------------------------------
// a minimal emulation of nx.Graph's ordered dict-of-dicts
// ---------------------------------------------------------------------------------------------------
struct NxEmu {
  int maxdeg;
  std::vector<int32_t> adj;    // [n x maxdeg], insertion order
  std::vector<uint8_t> deg;
  std::vector<uint8_t> alive;
  int64_t n = 0;
  explicit NxEmu(int md) : maxdeg(md) {}
  int32_t add_node() {
    adj.resize(adj.size() + maxdeg, -1);
    deg.push_back(0);
    alive.push_back(1);
    return (int32_t)n++;
  }
  void reserve(int64_t nn) { adj.reserve((size_t)nn * maxdeg); deg.reserve(nn); alive.reserve(nn); }
  bool has_edge(int32_t u, int32_t v) const {
    for (int j = 0; j < deg[u]; ++j) if (adj[(size_t)u * maxdeg + j] == v) return true;
    return false;
  }
  int add_edge(int32_t u, int32_t v) {
    if (has_edge(u, v)) return 0;
    if (deg[u] >= maxdeg || deg[v] >= maxdeg) return 1;
    adj[(size_t)u * maxdeg + deg[u]++] = v;
    if (u != v) adj[(size_t)v * maxdeg + deg[v]++] = u;
    return 0;
  }
  void remove_node(int32_t x) {
    for (int j = 0; j < deg[x]; ++j) {
      int32_t y = adj[(size_t)x * maxdeg + j];
      if (y == x) continue;
      int32_t* row = &adj[(size_t)y * maxdeg];
      int k = 0;
      for (int i = 0; i < deg[y]; ++i) if (row[i] != x) row[k++] = row[i];
      deg[y] = (uint8_t)k;
    }
    deg[x] = 0;
    alive[x] = 0;
  }
};

// compact the emulated graph into a gds_hostgraph: node ids = rank among alive nodes in insertion order;
// edges as G.edges() yields them (reportviews.py: for n in nodes, for nbr in adj[n], skip nbrs already done)
static void finish(const NxEmu& G, const std::vector<int32_t>& labels, int label_dim, const std::vector<double>* pos,
                   gds_hostgraph& out) {
  std::vector<int32_t> newid((size_t)G.n, -1);
  int64_t V = 0;
  for (int64_t x = 0; x < G.n; ++x) if (G.alive[x]) newid[x] = (int32_t)V++;
  out.V = V;
  out.label_dim = label_dim;
  out.labels.resize((size_t)V * label_dim);
  if (pos) out.pos.resize((size_t)V * 2);
  out.adj_ptr.assign((size_t)V + 1, 0);
  for (int64_t x = 0; x < G.n; ++x) {
    if (!G.alive[x]) continue;
    int32_t id = newid[x];
    for (int d = 0; d < label_dim; ++d) out.labels[(size_t)id * label_dim + d] = labels[(size_t)x * label_dim + d];
    if (pos) { out.pos[(size_t)id * 2] = (*pos)[(size_t)x * 2]; out.pos[(size_t)id * 2 + 1] = (*pos)[(size_t)x * 2 + 1]; }
    out.adj_ptr[id + 1] = out.adj_ptr[id] + G.deg[x];
  }
  out.adj_nbr.resize((size_t)out.adj_ptr[V]);
  int64_t E = 0;
  for (int64_t x = 0; x < G.n; ++x) {
    if (!G.alive[x]) continue;
    int32_t id = newid[x];
    for (int j = 0; j < G.deg[x]; ++j) {
      int32_t nb = newid[G.adj[(size_t)x * G.maxdeg + j]];
      out.adj_nbr[(size_t)out.adj_ptr[id] + j] = nb;
      if (nb >= id) ++E;  // nbr not yet exhausted as an outer node
    }
  }
  out.E = E;
  out.eu.resize((size_t)E); out.ev.resize((size_t)E);
  int64_t e = 0;
  for (int64_t id = 0; id < V; ++id)
    for (int64_t a = out.adj_ptr[id]; a < out.adj_ptr[id + 1]; ++a) {
      int32_t nb = out.adj_nbr[(size_t)a];
      if (nb >= id) { out.eu[(size_t)e] = (int32_t)id; out.ev[(size_t)e] = nb; ++e; }
    }
}

// ---------------------------------------------------------------------------------------------------
// lattices
// ---------------------------------------------------------------------------------------------------
static int gen_grid2d(int64_t m, int64_t n, gds_hostgraph& out) {
  // lattice.py grid_2d_graph: nodes (i,j) row-major; edges ((i,j),(i-1,j)) for i>=1 for j; then ((i,j),(i,j-1))
  NxEmu G(4);
  G.reserve(m * n);
  std::vector<int32_t> labels((size_t)m * n * 2);
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < n; ++j) {
      int32_t id = G.add_node();
      labels[(size_t)id * 2] = (int32_t)i; labels[(size_t)id * 2 + 1] = (int32_t)j;
    }
  for (int64_t i = 1; i < m; ++i)
    for (int64_t j = 0; j < n; ++j) if (G.add_edge((int32_t)(i * n + j), (int32_t)((i - 1) * n + j))) return 1;
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 1; j < n; ++j) if (G.add_edge((int32_t)(i * n + j), (int32_t)(i * n + j - 1))) return 1;
  finish(G, labels, 2, nullptr, out);
  return 0;
}

static int gen_grid3d(int64_t d0, int64_t d1, int64_t d2, gds_hostgraph& out) {
  // lattice.py grid_graph(dim=[d0,d1,d2]): iterated cartesian_product + relabel(flatten).  Resulting labels
  // (a,b,c) have a<d2, b<d1, c<d0, nodes in lexicographic order, and G.edges() yields per node its +a, +b, +c
  // neighbour (checked against networkx in tests/test_lattice.py).  No faces in 3-D, so the
  // adjacency insertion order is not needed: emit the closed form directly (10^9 edges: no adjacency is built).
  int64_t A = d2, B = d1, C = d0;
  int64_t V = A * B * C;
  if (V >= (int64_t(1) << 31) - 64) return 2;
  out.V = V; out.label_dim = 3;
  out.labels.resize((size_t)V * 3);
  int64_t E = (A - 1) * B * C + A * (B - 1) * C + A * B * (C - 1);
  out.E = E;
  out.eu.resize((size_t)E); out.ev.resize((size_t)E);
  int64_t e = 0, id = 0;
  for (int64_t a = 0; a < A; ++a)
    for (int64_t b = 0; b < B; ++b)
      for (int64_t c = 0; c < C; ++c, ++id) {
        out.labels[(size_t)id * 3] = (int32_t)a; out.labels[(size_t)id * 3 + 1] = (int32_t)b; out.labels[(size_t)id * 3 + 2] = (int32_t)c;
        if (a + 1 < A) { out.eu[(size_t)e] = (int32_t)id; out.ev[(size_t)e] = (int32_t)(id + B * C); ++e; }
        if (b + 1 < B) { out.eu[(size_t)e] = (int32_t)id; out.ev[(size_t)e] = (int32_t)(id + C); ++e; }
        if (c + 1 < C) { out.eu[(size_t)e] = (int32_t)id; out.ev[(size_t)e] = (int32_t)(id + 1); ++e; }
      }
  return 0;
}

// label -> emulated id with creation on first mention (add_edges_from adds u, then v)
struct LabelMap {
  int64_t ni, nj;
  std::vector<int32_t> id;
  std::vector<int32_t>& labels;
  NxEmu& G;
  LabelMap(int64_t ni_, int64_t nj_, std::vector<int32_t>& l, NxEmu& g) : ni(ni_), nj(nj_), id((size_t)ni_ * nj_, -1), labels(l), G(g) {}
  int32_t get(int64_t i, int64_t j) {
    int32_t& r = id[(size_t)i * nj + j];
    if (r < 0) {
      r = G.add_node();
      labels.push_back((int32_t)i); labels.push_back((int32_t)j);
    }
    return r;
  }
  int32_t find(int64_t i, int64_t j) const { return id[(size_t)i * nj + j]; }
};

static int gen_triangular(int64_t m, int64_t n, gds_hostgraph& out) {
  if (m <= 0 || n <= 0) return 2;
  int64_t N = (n + 1) / 2;
  NxEmu G(6);
  std::vector<int32_t> labels;
  G.reserve((N + 1) * (m + 1));
  LabelMap L(N + 1, m + 1, labels, G);
  auto edge = [&](int64_t i0, int64_t j0, int64_t i1, int64_t j1) { int32_t u = L.get(i0, j0); int32_t v = L.get(i1, j1); return G.add_edge(u, v); };
  // lattice.py triangular_lattice_graph
  for (int64_t j = 0; j <= m; ++j) for (int64_t i = 0; i < N; ++i) if (edge(i, j, i + 1, j)) return 1;
  for (int64_t j = 0; j < m; ++j) for (int64_t i = 0; i <= N; ++i) if (edge(i, j, i, j + 1)) return 1;
  for (int64_t j = 1; j < m; j += 2) for (int64_t i = 0; i < N; ++i) if (edge(i, j, i + 1, j + 1)) return 1;
  for (int64_t j = 0; j < m; j += 2) for (int64_t i = 0; i < N; ++i) if (edge(i + 1, j, i, j + 1)) return 1;
  if (n % 2) for (int64_t j = 1; j <= m; j += 2) { int32_t x = L.find(N, j); if (x >= 0) G.remove_node(x); }
  std::vector<double> pos((size_t)G.n * 2);
  const double h = std::sqrt(3.0) / 2.0;
  for (int64_t x = 0; x < G.n; ++x) {
    int64_t i = labels[(size_t)x * 2], j = labels[(size_t)x * 2 + 1];
    pos[(size_t)x * 2] = 0.5 * (double)(j % 2) + (double)i;
    pos[(size_t)x * 2 + 1] = h * (double)j;
  }
  finish(G, labels, 2, &pos, out);
  return 0;
}

static int gen_hexagonal(int64_t m, int64_t n, gds_hostgraph& out) {
  if (m <= 0 || n <= 0) return 2;
  int64_t M = 2 * m;
  NxEmu G(3);
  std::vector<int32_t> labels;
  G.reserve((n + 1) * (M + 2));
  LabelMap L(n + 1, M + 2, labels, G);
  auto edge = [&](int64_t i0, int64_t j0, int64_t i1, int64_t j1) { int32_t u = L.get(i0, j0); int32_t v = L.get(i1, j1); return G.add_edge(u, v); };
  // lattice.py hexagonal_lattice_graph
  for (int64_t i = 0; i <= n; ++i) for (int64_t j = 0; j <= M; ++j) if (edge(i, j, i, j + 1)) return 1;
  for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < M + 2; ++j) if (i % 2 == j % 2) if (edge(i, j, i + 1, j)) return 1;
  G.remove_node(L.find(0, M + 1));
  G.remove_node(L.find(n, (M + 1) * (n % 2)));
  std::vector<double> pos((size_t)G.n * 2);
  const double h = std::sqrt(3.0) / 2.0;
  for (int64_t x = 0; x < G.n; ++x) {
    int64_t i = labels[(size_t)x * 2], j = labels[(size_t)x * 2 + 1];
    pos[(size_t)x * 2] = 0.5 + (double)i + (double)(i / 2) + (double)(j % 2) * ((double)(i % 2) - 0.5);
    pos[(size_t)x * 2 + 1] = h * (double)j;
  }
  finish(G, labels, 2, &pos, out);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// random geometric graph: networkx/generators/geometric.py random_geometric_graph(n, r, seed=int)
//   pos[v] = [rng.random(), rng.random()] from Python's random.Random(seed)  (MT19937, 53-bit doubles)
//   edges  = sorted (u < v) pairs within distance r (the KD-tree query_pairs set, sorted by networkx)
// ---------------------------------------------------------------------------------------------------
struct PyRandom {
  uint32_t mt[624];
  int idx;
  void init_genrand(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  // CPython random_seed(): init_by_array over the 32-bit digits of abs(seed)
  explicit PyRandom(uint64_t seed) {
    uint32_t key[2];
    int klen = 1;
    key[0] = (uint32_t)(seed & 0xffffffffu);
    if (seed >> 32) { key[1] = (uint32_t)(seed >> 32); klen = 2; }
    init_genrand(19650218u);
    int i = 1, j = 0;
    int k = (624 > klen) ? 624 : klen;
    for (; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      ++i; ++j;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
      if (j >= klen) j = 0;
    }
    for (k = 623; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      ++i;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    idx = 624;
  }
  uint32_t next32() {
    if (idx >= 624) {
      for (int kk = 0; kk < 624; ++kk) {
        uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
        mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
  }
  double random() {
    uint32_t a = next32() >> 5, b = next32() >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
};

static int gen_rgg(int64_t n, double r, uint64_t seed, gds_hostgraph& out) {
  if (n <= 0 || !(r > 0)) return 2;
  PyRandom rng(seed);
  out.V = n; out.label_dim = 0;
  out.pos.resize((size_t)n * 2);
  for (int64_t v = 0; v < n; ++v) { out.pos[(size_t)v * 2] = rng.random(); out.pos[(size_t)v * 2 + 1] = rng.random(); }
  // cell binning with cell size >= r
  int64_t nc = (int64_t)std::floor(1.0 / r);
  if (nc < 1) nc = 1;
  if (nc > 8192) nc = 8192;
  double cs = 1.0 / (double)nc;
  std::vector<int64_t> cptr((size_t)nc * nc + 1, 0);
  std::vector<int32_t> cell_of((size_t)n);
  for (int64_t v = 0; v < n; ++v) {
    int64_t cx = std::min<int64_t>(nc - 1, (int64_t)(out.pos[(size_t)v * 2] / cs));
    int64_t cy = std::min<int64_t>(nc - 1, (int64_t)(out.pos[(size_t)v * 2 + 1] / cs));
    cell_of[(size_t)v] = (int32_t)(cx * nc + cy);
    cptr[(size_t)cell_of[(size_t)v] + 1]++;
  }
  for (size_t c = 0; c < (size_t)nc * nc; ++c) cptr[c + 1] += cptr[c];
  std::vector<int32_t> members((size_t)n);
  {
    std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t v = 0; v < n; ++v) members[(size_t)fill[(size_t)cell_of[(size_t)v]]++] = (int32_t)v;  // ascending v inside a cell
  }
  const double r2 = r * r;
  // pass 1: count higher neighbours per node, pass 2: fill, then sort each node's list => lexicographic (u,v)
  std::vector<int64_t> ptr((size_t)n + 1, 0);
  auto for_pairs = [&](auto&& visit) {
    for (int64_t u = 0; u < n; ++u) {
      double ux = out.pos[(size_t)u * 2], uy = out.pos[(size_t)u * 2 + 1];
      int64_t cx = cell_of[(size_t)u] / nc, cy = cell_of[(size_t)u] % nc;
      for (int64_t dx = -1; dx <= 1; ++dx) {
        int64_t x = cx + dx;
        if (x < 0 || x >= nc) continue;
        for (int64_t dy = -1; dy <= 1; ++dy) {
          int64_t y = cy + dy;
          if (y < 0 || y >= nc) continue;
          int64_t c = x * nc + y;
          for (int64_t a = cptr[(size_t)c]; a < cptr[(size_t)c + 1]; ++a) {
            int32_t v = members[(size_t)a];
            if (v <= u) continue;
            double ddx = out.pos[(size_t)v * 2] - ux, ddy = out.pos[(size_t)v * 2 + 1] - uy;
            // scipy cKDTree query_pairs(r, p=2): squared distance <= r^2
            if (ddx * ddx + ddy * ddy <= r2) visit(u, v);
          }
        }
      }
    }
  };
  for_pairs([&](int64_t u, int32_t) { ptr[(size_t)u + 1]++; });
  for (int64_t u = 0; u < n; ++u) ptr[(size_t)u + 1] += ptr[(size_t)u];
  int64_t E = ptr[(size_t)n];
  if (E >= (int64_t(1) << 30)) return 2;
  out.E = E;
  out.eu.resize((size_t)E); out.ev.resize((size_t)E);
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for_pairs([&](int64_t u, int32_t v) { int64_t a = fill[(size_t)u]++; out.eu[(size_t)a] = (int32_t)u; out.ev[(size_t)a] = v; });
  }
  for (int64_t u = 0; u < n; ++u) std::sort(out.ev.begin() + ptr[(size_t)u], out.ev.begin() + ptr[(size_t)u + 1]);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// planar face walk (gds/utils/graph/generators.py:444-511)
// ---------------------------------------------------------------------------------------------------
static inline double turn_angle(const double* pos, int32_t a, int32_t b, int32_t c) {
  const double two_pi = 2.0 * M_PI;
  double x1 = pos[(size_t)b * 2] - pos[(size_t)a * 2], y1 = pos[(size_t)b * 2 + 1] - pos[(size_t)a * 2 + 1];
  double x2 = pos[(size_t)c * 2] - pos[(size_t)b * 2], y2 = pos[(size_t)c * 2 + 1] - pos[(size_t)b * 2 + 1];
  double theta = std::atan2(y2, x2) - std::atan2(y1, x1);
  double magn = std::fmod(std::fabs(theta), two_pi);
  theta = (theta < 0) ? (two_pi - magn) : magn;   // generators.py:454-459
  return (theta > M_PI) ? (theta - two_pi) : theta;
}

static bool point_in_closed_polygon(const double* pos, const int32_t* poly, int64_t k, double x, double y) {
  bool inside = false;
  for (int64_t i = 0; i < k; ++i) {
    double x1 = pos[(size_t)poly[i] * 2], y1 = pos[(size_t)poly[i] * 2 + 1];
    double x2 = pos[(size_t)poly[(i + 1) % k] * 2], y2 = pos[(size_t)poly[(i + 1) % k] * 2 + 1];
    double cross = (x2 - x1) * (y - y1) - (y2 - y1) * (x - x1);
    if (std::fabs(cross) <= 1e-12 * std::max(1.0, std::fabs(x2 - x1) + std::fabs(y2 - y1)) &&
        std::min(x1, x2) - 1e-12 <= x && x <= std::max(x1, x2) + 1e-12 && std::min(y1, y2) - 1e-12 <= y && y <= std::max(y1, y2) + 1e-12)
      return true;
    if ((y1 > y) != (y2 > y)) {
      double xin = x1 + (y - y1) * (x2 - x1) / (y2 - y1);
      if (xin > x) inside = !inside;
    }
  }
  return inside;
}

static int walk_faces(gds_hostgraph& g) {
  const int64_t V = g.V, E = g.E;
  if (g.pos.empty()) GDS_FAIL("face walk needs node positions");
  const double* pos = g.pos.data();
  std::vector<uint8_t> seen(g.adj_nbr.size(), 0);  // half-edge (a -> adj_nbr[slot]) seen
  auto slot_of = [&](int32_t a, int32_t b) -> int64_t {
    for (int64_t s = g.adj_ptr[(size_t)a]; s < g.adj_ptr[(size_t)a + 1]; ++s) if (g.adj_nbr[(size_t)s] == b) return s;
    return -1;
  };
  std::vector<int64_t> fptr(1, 0);
  std::vector<int32_t> fnodes;
  std::vector<int32_t> path;
  for (int64_t e = 0; e < E; ++e) {
    for (int dir = 0; dir < 2; ++dir) {
      int32_t tail = dir ? g.ev[(size_t)e] : g.eu[(size_t)e], head = dir ? g.eu[(size_t)e] : g.ev[(size_t)e];
      int64_t s0 = slot_of(tail, head);
      if (s0 < 0) GDS_FAIL("edge missing from adjacency");
      if (seen[(size_t)s0]) continue;
      path.clear();
      path.push_back(head);
      bool ok = true;
      int64_t guard = 0;
      while (true) {
        int32_t best = -1;
        double best_angle = -std::numeric_limits<double>::infinity();
        for (int64_t s = g.adj_ptr[(size_t)head]; s < g.adj_ptr[(size_t)head + 1]; ++s) {
          int32_t nd = g.adj_nbr[(size_t)s];
          if (nd == tail) continue;
          double ang = turn_angle(pos, tail, head, nd);
          if (ang > best_angle) { best_angle = ang; best = nd; }
        }
        if (best < 0) { ok = false; break; }   // dead end: no face (generators.py:483-485)
        if (best == path[0]) break;
        path.push_back(best);
        tail = head; head = best;
        if (++guard > 2 * E + 2) GDS_FAIL("face walk does not close (graph is not plane-embedded by its positions)");
      }
      if (!ok) continue;
      int64_t k = (int64_t)path.size();
      fnodes.insert(fnodes.end(), path.begin(), path.end());
      fptr.push_back((int64_t)fnodes.size());
      for (int64_t i = 0; i < k; ++i) {
        int64_t s = slot_of(path[(size_t)((i + k - 1) % k)], path[(size_t)i]);
        if (s >= 0) seen[(size_t)s] = 1;
      }
    }
  }
  // outer face (generators.py:498-506): first face whose closed polygon touches every node.  A clockwise face whose
  // bounding box covers all nodes is the unbounded face; a counter-clockwise candidate gets the full point test.
  double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
  for (int64_t v = 0; v < V; ++v) {
    bx0 = std::min(bx0, pos[(size_t)v * 2]); bx1 = std::max(bx1, pos[(size_t)v * 2]);
    by0 = std::min(by0, pos[(size_t)v * 2 + 1]); by1 = std::max(by1, pos[(size_t)v * 2 + 1]);
  }
  int64_t nF = (int64_t)fptr.size() - 1, outer = -1;
  for (int64_t f = 0; f < nF && outer < 0; ++f) {
    const int32_t* poly = &fnodes[(size_t)fptr[(size_t)f]];
    int64_t k = fptr[(size_t)f + 1] - fptr[(size_t)f];
    double fx0 = 1e300, fx1 = -1e300, fy0 = 1e300, fy1 = -1e300, area = 0;
    for (int64_t i = 0; i < k; ++i) {
      double x = pos[(size_t)poly[i] * 2], y = pos[(size_t)poly[i] * 2 + 1];
      double xn = pos[(size_t)poly[(i + 1) % k] * 2], yn = pos[(size_t)poly[(i + 1) % k] * 2 + 1];
      fx0 = std::min(fx0, x); fx1 = std::max(fx1, x); fy0 = std::min(fy0, y); fy1 = std::max(fy1, y);
      area += x * yn - xn * y;
    }
    if (fx0 > bx0 || fx1 < bx1 || fy0 > by0 || fy1 < by1) continue;
    if (k >= 3 && area < 0) { outer = f; break; }
    bool all = true;
    for (int64_t v = 0; v < V && all; ++v) all = point_in_closed_polygon(pos, poly, k, pos[(size_t)v * 2], pos[(size_t)v * 2 + 1]);
    if (all) outer = f;
  }
  g.face_ptr.assign(1, 0);
  g.face_nodes.clear();
  for (int64_t f = 0; f < nF; ++f) {
    if (f == outer) continue;
    g.face_nodes.insert(g.face_nodes.end(), fnodes.begin() + fptr[(size_t)f], fnodes.begin() + fptr[(size_t)f + 1]);
    g.face_ptr.push_back((int64_t)g.face_nodes.size());
  }
  g.F = (int64_t)g.face_ptr.size() - 1;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_hostgraph_lattice(int32_t kind, int64_t a, int64_t b, int64_t c, double r, int32_t with_faces,
                                     gds_hostgraph** out) {
  if (!out) GDS_FAIL("null out");
  gds_hostgraph* g = new gds_hostgraph();
  int rc = 0;
  switch (kind) {
    case 0: rc = gen_grid2d(a, b, *g); break;
    case 1: rc = gen_grid3d(a, b, c, *g); break;
    case 2: rc = gen_triangular(a, b, *g); break;
    case 3: rc = gen_hexagonal(a, b, *g); break;
    case 4: rc = gen_rgg(a, r, (uint64_t)b, *g); break;
    default: rc = 3;
  }
  if (rc) { delete g; GDS_FAIL(rc == 1 ? "degree overflow in lattice emulation" : (rc == 2 ? "bad lattice size" : "unknown lattice kind")); }
  g->face_ptr.assign(1, 0);
  if (with_faces && (kind == 2 || kind == 3)) {
    rc = walk_faces(*g);
    if (rc) { delete g; return rc; }
  }
  *out = g;
  return 0;
}

extern "C" int gds_hostgraph_from_graph(int64_t V, int64_t E, const int32_t* eu, const int32_t* ev, const int64_t* adj_ptr,
                                        const int32_t* adj_nbr, const double* pos, gds_hostgraph** out) {
  if (!out || !eu || !ev || !adj_ptr || !adj_nbr || !pos) GDS_FAIL("bad argument");
  gds_hostgraph* g = new gds_hostgraph();
  g->V = V; g->E = E;
  g->eu.assign(eu, eu + E); g->ev.assign(ev, ev + E);
  g->adj_ptr.assign(adj_ptr, adj_ptr + V + 1);
  g->adj_nbr.assign(adj_nbr, adj_nbr + adj_ptr[V]);
  g->pos.assign(pos, pos + 2 * V);
  int rc = walk_faces(*g);
  if (rc) { delete g; return rc; }
  *out = g;
  return 0;
}

extern "C" int gds_hostgraph_sizes(gds_hostgraph* g, int64_t* V, int64_t* E, int64_t* F, int64_t* face_nnz,
                                   int32_t* label_dim, int32_t* has_pos) {
  if (!g) GDS_FAIL("null graph");
  if (V) *V = g->V;
  if (E) *E = g->E;
  if (F) *F = g->F;
  if (face_nnz) *face_nnz = (int64_t)g->face_nodes.size();
  if (label_dim) *label_dim = g->label_dim;
  if (has_pos) *has_pos = g->pos.empty() ? 0 : 1;
  return 0;
}

extern "C" int gds_hostgraph_copy(gds_hostgraph* g, int32_t* labels, double* pos, int32_t* eu, int32_t* ev,
                                  int64_t* face_ptr, int32_t* face_nodes) {
  if (!g) GDS_FAIL("null graph");
  if (labels) std::copy(g->labels.begin(), g->labels.end(), labels);
  if (pos) std::copy(g->pos.begin(), g->pos.end(), pos);
  if (eu) std::copy(g->eu.begin(), g->eu.end(), eu);
  if (ev) std::copy(g->ev.begin(), g->ev.end(), ev);
  if (face_ptr) std::copy(g->face_ptr.begin(), g->face_ptr.end(), face_ptr);
  if (face_nodes) std::copy(g->face_nodes.begin(), g->face_nodes.end(), face_nodes);
  return 0;
}

extern "C" int gds_hostgraph_destroy(gds_hostgraph* g) {
  delete g;
  return 0;
}

// lower a native host graph straight to the device (no copy through the caller)
extern "C" int gds_graph_create_from_host(gds_ctx* ctx, gds_hostgraph* g, const double* weight, gds_graph** out) {
  if (!g) GDS_FAIL("null host graph");
  return gds_graph_create(ctx, g->V, g->E, g->eu.data(), g->ev.data(), weight, g->F, g->face_ptr.data(),
                          g->face_nodes.data(), out);
}
