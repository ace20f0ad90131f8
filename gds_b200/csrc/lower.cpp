// Host-side operator assembly: edge list (+ faces) -> incidence index maps.
// Replaces the Python-level assembly of gds/gds.py:16-55,105-109,197,227-239,253-262 (dense / O(E^2) there).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "internal.h"

int lower_graph_host(int64_t V, int64_t E, const int32_t* eu, const int32_t* ev, const double* w, int64_t F,
                     const int64_t* face_ptr, const int32_t* face_nodes, HostGraph& g) {
  if (V < 0 || E < 0 || F < 0) GDS_FAIL("negative size");
  if (V >= (int64_t(1) << 31) - 64 || E >= (int64_t(1) << 30)) GDS_FAIL("graph too large for int32 indices");
  g.V = V; g.E = E; g.F = F;
  g.eu.assign(eu, eu + E);
  g.ev.assign(ev, ev + E);
  g.w.resize(E); g.omega.resize(E);
  g.unit_weights = true;
  for (int64_t e = 0; e < E; ++e) {
    if (eu[e] < 0 || eu[e] >= V || ev[e] < 0 || ev[e] >= V) GDS_FAIL("edge endpoint out of range");
    double we = w ? w[e] : 1.0;
    g.w[e] = we;
    g.omega[e] = std::sqrt(we);  // gds.py:109
    if (we != 1.0) g.unit_weights = false;
  }
  // ---- B1 rows: counting sort by node; walking edges in index order keeps each row sorted by edge id
  g.n_ptr.assign(V + 1, 0);
  for (int64_t e = 0; e < E; ++e) { g.n_ptr[eu[e] + 1]++; g.n_ptr[ev[e] + 1]++; }
  for (int64_t x = 0; x < V; ++x) g.n_ptr[x + 1] += g.n_ptr[x];
  g.n_edge.resize(2 * E); g.n_nbr.resize(2 * E); g.n_sign.resize(2 * E);
  {
    std::vector<int64_t> fill(g.n_ptr.begin(), g.n_ptr.end() - 1);
    for (int64_t e = 0; e < E; ++e) {
      int64_t a = fill[eu[e]]++;
      g.n_edge[a] = (int32_t)e; g.n_nbr[a] = ev[e]; g.n_sign[a] = -1;  // B[u,e] = -sqrt(w)
      int64_t b = fill[ev[e]]++;
      g.n_edge[b] = (int32_t)e; g.n_nbr[b] = eu[e]; g.n_sign[b] = +1;  // B[v,e] = +sqrt(w)
    }
  }
  // ---- 1/(weighted degree - 1), inf -> 0   (gds.py:253-257, utils/graph/common.py:5-10)
  g.n_dinv.resize(V);
  for (int64_t x = 0; x < V; ++x) {
    double deg = 0.0;
    for (int64_t a = g.n_ptr[x]; a < g.n_ptr[x + 1]; ++a) deg += g.w[g.n_edge[a]];
    double d = deg - 1.0;
    g.n_dinv[x] = (d == 0.0) ? 0.0 : 1.0 / d;
  }
  // ---- B2 rows (gds.py:227-236): every edge with BOTH endpoints on the face; + if its stored orientation is a
  //      consecutive pair of the face traversal, - otherwise
  g.f_ptr.assign(F + 1, 0);
  g.f_edge.clear(); g.f_sign.clear();
  if (F > 0) {
    std::vector<int64_t> stamp(V, -1);
    std::vector<int64_t> fwd_stamp(E, -1);
    std::vector<std::pair<int32_t, int8_t>> row;
    for (int64_t f = 0; f < F; ++f) {
      int64_t p0 = face_ptr[f], p1 = face_ptr[f + 1];
      int64_t k = p1 - p0;
      for (int64_t i = 0; i < k; ++i) {
        int32_t nd = face_nodes[p0 + i];
        if (nd < 0 || nd >= V) GDS_FAIL("face node out of range");
        stamp[nd] = f;
      }
      // forward pairs (f[i-1], f[i]) and (f[k-1], f[0])
      for (int64_t i = 0; i < k; ++i) {
        int32_t a = face_nodes[p0 + i], b = face_nodes[p0 + (i + 1) % k];
        for (int64_t s = g.n_ptr[a]; s < g.n_ptr[a + 1]; ++s)
          if (g.n_nbr[s] == b && g.n_sign[s] == -1) fwd_stamp[g.n_edge[s]] = f;  // stored as (a -> b)
      }
      row.clear();
      for (int64_t i = 0; i < k; ++i) {
        int32_t a = face_nodes[p0 + i];
        for (int64_t s = g.n_ptr[a]; s < g.n_ptr[a + 1]; ++s) {
          if (g.n_sign[s] != -1) continue;  // visit each edge from its tail only
          int32_t b = g.n_nbr[s];
          if (stamp[b] != f) continue;
          int32_t e = g.n_edge[s];
          row.emplace_back(e, (int8_t)(fwd_stamp[e] == f ? +1 : -1));
        }
      }
      std::sort(row.begin(), row.end());
      row.erase(std::unique(row.begin(), row.end()), row.end());  // a node listed twice on a face
      for (auto& pr : row) { g.f_edge.push_back(pr.first); g.f_sign.push_back(pr.second); }
      g.f_ptr[f + 1] = (int64_t)g.f_edge.size();
    }
  }
  // ---- B2 columns: edge -> faces (transpose; rows stay sorted by face id)
  g.ef_ptr.assign(E + 1, 0);
  int64_t nnz2 = (int64_t)g.f_edge.size();
  for (int64_t a = 0; a < nnz2; ++a) g.ef_ptr[g.f_edge[a] + 1]++;
  for (int64_t e = 0; e < E; ++e) g.ef_ptr[e + 1] += g.ef_ptr[e];
  g.ef_face.resize(nnz2); g.ef_sign.resize(nnz2);
  {
    std::vector<int64_t> fill(g.ef_ptr.begin(), g.ef_ptr.end() - 1);
    for (int64_t f = 0; f < F; ++f)
      for (int64_t a = g.f_ptr[f]; a < g.f_ptr[f + 1]; ++a) {
        int64_t s = fill[g.f_edge[a]]++;
        g.ef_face[s] = (int32_t)f; g.ef_sign[s] = g.f_sign[a];
      }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Aggregation for the multigrid preconditioner of the pressure / Leray solves (include/gds_b200.h, gds_mg_*).
// Input: a symmetric matrix in CSR (a masked graph Laplacian or a Galerkin coarse operator).  j is a STRONG neighbour of i
// when |a_ij| >= theta * max_{k != i} |a_ik| (i != j, a_ij != 0; 0 <= theta <= 1, so the largest off-diagonal entry of a
// row is always strong whatever the row's degree).  Three sweeps in index order, so the result is a pure
// function of the matrix:  (1) a node whose strong neighbours are all free founds an aggregate with them;  (2) a node left
// over joins the aggregate of its most strongly connected aggregated neighbour (as it stood after sweep 1);  (3) what is
// still free founds aggregates with its free strong neighbours.  Rows without a strong neighbour (Dirichlet rows kept as
// identity rows, isolated nodes) stay outside: agg = -1, they have no coarse representation.
extern "C" int gds_mg_aggregate(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, double theta,
                                int32_t* agg, int64_t* n_agg) {
  if (n < 0 || !rowptr || (n > 0 && (!col || !val)) || !agg || !n_agg || !(theta >= 0.0 && theta <= 1.0)) GDS_FAIL("bad argument");
  if (n >= (int64_t(1) << 31) - 64) GDS_FAIL("too many rows for int32 aggregate ids");
  std::vector<double> rowmax((size_t)n, 0.0);     // largest off-diagonal magnitude of every row
  for (int64_t i = 0; i < n; ++i) {
    if (rowptr[i + 1] < rowptr[i]) GDS_FAIL("row pointers must not decrease");
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      if (col[k] < 0 || col[k] >= n) GDS_FAIL("column index out of range");
      if (col[k] != i) rowmax[(size_t)i] = std::max(rowmax[(size_t)i], std::fabs(val[k]));
    }
  }
  auto strong = [&](int64_t i, int64_t k) {
    if (col[k] == i || val[k] == 0.0) return false;
    return std::fabs(val[k]) >= theta * rowmax[(size_t)i];
  };
  std::fill(agg, agg + n, (int32_t)-1);
  std::vector<uint8_t> has_strong((size_t)n, 0);
  int32_t na = 0;
  for (int64_t i = 0; i < n; ++i) {                       // sweep 1
    bool any = false, all_free = agg[i] < 0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (strong(i, k)) { any = true; if (agg[col[k]] >= 0) all_free = false; }
    has_strong[(size_t)i] = any;
    if (!any || !all_free) continue;
    agg[i] = na;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (strong(i, k)) agg[col[k]] = na;
    ++na;
  }
  {                                                         // sweep 2 (reads the result of sweep 1 only)
    std::vector<int32_t> joined((size_t)n, -1);
    for (int64_t i = 0; i < n; ++i) {
      if (agg[i] >= 0 || !has_strong[(size_t)i]) continue;
      double best = -1.0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
        if (strong(i, k) && agg[col[k]] >= 0 && std::fabs(val[k]) > best) { best = std::fabs(val[k]); joined[(size_t)i] = agg[col[k]]; }
    }
    for (int64_t i = 0; i < n; ++i)
      if (joined[(size_t)i] >= 0) agg[i] = joined[(size_t)i];
  }
  for (int64_t i = 0; i < n; ++i) {                       // sweep 3
    if (agg[i] >= 0 || !has_strong[(size_t)i]) continue;
    agg[i] = na;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (strong(i, k) && agg[col[k]] < 0) agg[col[k]] = na;
    ++na;
  }
  *n_agg = na;
  return 0;
}
