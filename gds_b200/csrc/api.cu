// C-ABI implementation: contexts, device vectors, graph upload, pass resolution, systems and CUDA-graph stepping.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "internal.h"

static thread_local std::string g_last_error;
void gds_set_error(const std::string& msg) { g_last_error = msg; }

extern "C" const char* gds_last_error(void) { return g_last_error.c_str(); }
extern "C" int gds_abi_version(void) { return GDS_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_ctx_create(int device, gds_ctx** out) {
  if (!out) GDS_FAIL("null out");
  if (device == -1) {
    // host-only context: operator assembly and row-program validation without a GPU (index-map parity checks,
    // dry runs of the tracer).  It never computes: every call that would touch the device fails.
    gds_ctx* c = new gds_ctx();
    c->device = -1;
    *out = c;
    return 0;
  }
  int ndev = 0;
  GDS_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) GDS_FAIL("no such CUDA device");
  GDS_CUDA(cudaSetDevice(device));
  gds_ctx* c = new gds_ctx();
  c->device = device;
  // the context's stream carries the boundary rows / push / wait branch of an overlapped step and gets the higher
  // priority; the side stream carries the interior rows, whose grid would otherwise keep small late kernels waiting
  int prio_lo = 0, prio_hi = 0;
  GDS_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  GDS_CUDA(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio_hi));
  c->stream = c->own_stream;
  GDS_CUDA(cudaEventCreate(&c->ev0));
  GDS_CUDA(cudaEventCreate(&c->ev1));
  GDS_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio_lo));
  for (auto& e : c->ev_pool) GDS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = c;
  return 0;
}

#define GDS_NEED_DEVICE(c)                                                                  \
  do {                                                                                      \
    if ((c)->device < 0) GDS_FAIL("host-only context: no CUDA device behind this call");    \
  } while (0)

extern "C" int gds_ctx_destroy(gds_ctx* c) {
  if (!c) return 0;
  if (c->device < 0) { delete c; return 0; }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->reduce_scratch) cudaFree(c->reduce_scratch);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  for (auto& e : c->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(c->side_stream);
  cudaStreamDestroy(c->own_stream);
  delete c;
  return 0;
}

extern "C" int gds_ctx_set_stream(gds_ctx* c, void* s) {
  if (!c) GDS_FAIL("null ctx");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return 0;
}

extern "C" int gds_ctx_sync(gds_ctx* c) {
  if (!c) GDS_FAIL("null ctx");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  GDS_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gds_ctx_timer_start(gds_ctx* c) {
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaEventRecord(c->ev0, c->stream));
  return 0;
}

extern "C" int gds_ctx_timer_stop(gds_ctx* c, double* ms_out) {
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaEventRecord(c->ev1, c->stream));
  GDS_CUDA(cudaEventSynchronize(c->ev1));
  float ms = 0;
  GDS_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  if (ms_out) *ms_out = (double)ms;
  return 0;
}

extern "C" int gds_ctx_counters(gds_ctx* c, int64_t* kl, int64_t* gr) {
  if (kl) *kl = c->kernel_launches;
  if (gr) *gr = c->graph_replays;
  return 0;
}

extern "C" int gds_ctx_flush_l2(gds_ctx* c, int64_t bytes) {
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  if (bytes > c->flush_bytes) {
    if (c->flush_buf) GDS_CUDA(cudaFree(c->flush_buf));
    c->flush_buf = nullptr;
    GDS_CUDA(cudaMalloc(&c->flush_buf, (size_t)bytes));
    c->flush_bytes = bytes;
  }
  GDS_CUDA(cudaMemsetAsync(c->flush_buf, 0, (size_t)bytes, c->stream));
  return 0;
}

// Measurement aid: device time of `reps` launches of a gather-free kernel that reads n_read and writes n_write
// float64 vectors of n elements (the stream mix of one RK stage); returns the achieved GB/s.
extern "C" int gds_ctx_bandwidth_probe(gds_ctx* c, int32_t n_read, int32_t n_write, int64_t n, int32_t reps, double* gbs_out) {
  if (!c || n_read < 0 || n_read > 8 || n_write < 0 || n_write > 4 || n < 2 || reps < 1) GDS_FAIL("bad argument");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  std::vector<double*> bufs((size_t)(n_read + n_write), nullptr);
  int rc = 0;
  for (auto& b : bufs) {
    if (cudaMalloc((void**)&b, (size_t)n * 8) != cudaSuccess) { rc = 1; break; }
    cudaMemsetAsync(b, 0, (size_t)n * 8, c->stream);
  }
  float ms = 0.f;
  if (!rc) {
    for (int i = 0; i < 2 && !rc; ++i) rc = launch_probe(c, bufs.data(), n_read, n_write, n);
    cudaEventRecord(c->ev0, c->stream);
    for (int i = 0; i < reps && !rc; ++i) rc = launch_probe(c, bufs.data(), n_read, n_write, n);
    cudaEventRecord(c->ev1, c->stream);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) rc = 1;
  }
  for (auto& b : bufs) cudaFree(b);
  if (rc) GDS_FAIL("probe failed (out of memory?)");
  if (gbs_out) *gbs_out = (double)(n_read + n_write) * (double)n * 8.0 * reps / (ms * 1e-3) / 1e9;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// device vectors and masks
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_buf_create(gds_ctx* c, int64_t n, gds_buf** out) {
  if (!c || !out || n < 0) GDS_FAIL("bad argument");
  if (c->device < 0) {  // shape only
    gds_buf* b = new gds_buf();
    b->ctx = c; b->n = n;
    *out = b;
    return 0;
  }
  GDS_CUDA(cudaSetDevice(c->device));
  gds_buf* b = new gds_buf();
  b->ctx = c;
  b->n = n;
  // +4 doubles: double4 stores of the aggregate pass never run past the allocation
  GDS_CUDA(cudaMalloc((void**)&b->d, (size_t)(n + 4) * sizeof(double)));
  GDS_CUDA(cudaMemsetAsync(b->d, 0, (size_t)(n + 4) * sizeof(double), c->stream));
  *out = b;
  return 0;
}

// a view of n values at a device address owned elsewhere (e.g. a slice of a system's state) usable as a pass operand
extern "C" int gds_buf_view(gds_ctx* c, void* dev_ptr, int64_t n, gds_buf** out) {
  if (!c || !out || !dev_ptr || n < 0) GDS_FAIL("bad argument");
  gds_buf* b = new gds_buf();
  b->ctx = c; b->d = (double*)dev_ptr; b->n = n; b->owned = false;
  *out = b;
  return 0;
}

extern "C" int gds_buf_destroy(gds_buf* b) {
  if (!b) return 0;
  if (b->ctx->device < 0 || !b->owned) { delete b; return 0; }
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  cudaFree(b->d);
  delete b;
  return 0;
}

extern "C" int gds_buf_upload(gds_buf* b, int64_t off, const double* host, int64_t n) {
  if (!b || off < 0 || n < 0 || off + n > b->n) GDS_FAIL("range outside buffer");
  if (b->ctx->device < 0) return 0;  // host-only dry run: nothing to hold the data
  GDS_CUDA(cudaSetDevice(b->ctx->device));
  GDS_CUDA(cudaMemcpyAsync(b->d + off, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, b->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(b->ctx->stream));
  return 0;
}

extern "C" int gds_buf_download(gds_buf* b, int64_t off, double* host, int64_t n) {
  if (!b || off < 0 || n < 0 || off + n > b->n) GDS_FAIL("range outside buffer");
  GDS_NEED_DEVICE(b->ctx);
  GDS_CUDA(cudaSetDevice(b->ctx->device));
  GDS_CUDA(cudaMemcpyAsync(host, b->d + off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, b->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(b->ctx->stream));
  return 0;
}

extern "C" int gds_buf_fill(gds_buf* b, int64_t off, int64_t n, double v) {
  if (!b || off < 0 || n < 0 || off + n > b->n) GDS_FAIL("range outside buffer");
  GDS_NEED_DEVICE(b->ctx);
  GDS_CUDA(cudaSetDevice(b->ctx->device));
  return launch_fill(b->ctx, b->d + off, n, v);
}

// scalar reduction of n values of a device vector (kind 0 sum, 1 max, 2 min): only the 8-byte result crosses PCIe
extern "C" int gds_buf_reduce(gds_buf* b, int64_t off, int64_t n, int32_t kind, double* out) {
  if (!b || !out || off < 0 || n < 0 || off + n > b->n) GDS_FAIL("range outside buffer");
  if (kind < 0 || kind > 2) GDS_FAIL("unknown reduction");
  gds_ctx* c = b->ctx;
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  if (!c->reduce_scratch) GDS_CUDA(cudaMalloc((void**)&c->reduce_scratch, 2048 * sizeof(double)));
  GDS_TRY(launch_reduce(c, b->d + off, n, kind, c->reduce_scratch));
  GDS_CUDA(cudaMemcpyAsync(out, c->reduce_scratch + 1184, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GDS_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gds_buf_devptr(gds_buf* b, void** dev_out, int64_t* n_out) {
  if (!b) GDS_FAIL("null buffer");
  if (dev_out) *dev_out = b->d;
  if (n_out) *n_out = b->n;
  return 0;
}

extern "C" int gds_mask_create(gds_ctx* c, int64_t n_rows, const int64_t* rows, int64_t n, gds_mask** out) {
  if (!c || !out || n_rows < 0 || n < 0) GDS_FAIL("bad argument");
  if (c->device >= 0) GDS_CUDA(cudaSetDevice(c->device));
  int64_t words = (n_rows + 31) / 32 + 1;
  std::vector<uint32_t> bits((size_t)words, 0u);
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = rows[i];
    if (r < 0 || r >= n_rows) GDS_FAIL("mask row out of range");
    bits[(size_t)(r >> 5)] |= (1u << (r & 31));
  }
  gds_mask* m = new gds_mask();
  m->ctx = c; m->n_rows = n_rows; m->n_set = n;
  if (c->device < 0) { *out = m; return 0; }
  GDS_CUDA(cudaMalloc((void**)&m->d, (size_t)words * 4));
  GDS_CUDA(cudaMemcpyAsync(m->d, bits.data(), (size_t)words * 4, cudaMemcpyHostToDevice, c->stream));
  GDS_CUDA(cudaStreamSynchronize(c->stream));
  *out = m;
  return 0;
}

extern "C" int gds_mask_destroy(gds_mask* m) {
  if (!m) return 0;
  if (m->ctx->device < 0) { delete m; return 0; }
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  cudaFree(m->d);
  delete m;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// graph upload
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int upload(gds_graph* g, const std::vector<T>& h, const T** dev) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  void* d = nullptr;
  GDS_CUDA(cudaMalloc(&d, bytes));
  if (!h.empty()) GDS_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  g->allocs.push_back(d);
  g->device_bytes += (int64_t)bytes;
  *dev = (const T*)d;
  return 0;
}

// CSR rows -> sliced-ELL offsets; slot(row, j) = soff[row/32]*32 + j*32 + row%32
static void sell_offsets(int64_t n_rows, const std::vector<int64_t>& ptr, std::vector<uint32_t>& soff) {
  int64_t n_slices = (n_rows + GDS_SLICE - 1) / GDS_SLICE;
  soff.assign((size_t)n_slices + 2, 0u);
  for (int64_t s = 0; s < n_slices; ++s) {
    int64_t wmax = 0;
    int64_t r1 = std::min(n_rows, (s + 1) * GDS_SLICE);
    for (int64_t r = s * GDS_SLICE; r < r1; ++r) wmax = std::max(wmax, ptr[r + 1] - ptr[r]);
    soff[s + 1] = soff[s] + (uint32_t)wmax;
  }
  soff[n_slices + 1] = soff[n_slices];
}

extern "C" int gds_graph_create(gds_ctx* c, int64_t V, int64_t E, const int32_t* eu, const int32_t* ev,
                                const double* w, int64_t F, const int64_t* face_ptr, const int32_t* face_nodes,
                                gds_graph** out) {
  if (!c || !out) GDS_FAIL("bad argument");
  if (c->device >= 0) GDS_CUDA(cudaSetDevice(c->device));
  gds_graph* g = new gds_graph();
  g->ctx = c;
  int rc = lower_graph_host(V, E, eu, ev, w, F, face_ptr, face_nodes, g->h);
  if (rc) { delete g; return rc; }
  if (c->device < 0) {  // host-only: index maps for export, no device arrays
    memset(&g->d, 0, sizeof(g->d));
    g->d.V = V; g->d.E = E; g->d.F = F;
    *out = g;
    return 0;
  }
  HostGraph& h = g->h;
  GraphDev& d = g->d;
  memset(&d, 0, sizeof(d));
  d.V = V; d.E = E; d.F = F;
  // node incidence
  {
    std::vector<uint32_t> soff;
    sell_offsets(V, h.n_ptr, soff);
    {  // near-uniform degrees (lattices): pad every slice to the widest one when that costs < 2 % more slots, so that a
       // kernel can compute a slice's position instead of loading it (one dependent load less before the gather);
       // padded slots point at the row itself and add +0.0
      const int64_t n_slices = (int64_t)soff.size() - 1;
      uint32_t maxw = 0;
      for (int64_t i = 0; i < n_slices; ++i) maxw = std::max(maxw, soff[i + 1] - soff[i]);
      const int64_t uniform = (int64_t)maxw * n_slices, have = soff[n_slices];
      if (maxw > 0 && uniform <= have + have / 50 && uniform < (1ll << 31) && !getenv("GDS_B200_NO_UNIW")) {
        for (int64_t i = 0; i <= n_slices; ++i) soff[i] = (uint32_t)(i * maxw);
        d.n_uniw = (int32_t)maxw;
      }
    }
    int64_t slots = (int64_t)soff[soff.size() - 1] * GDS_SLICE;
    const bool unitw = h.unit_weights;  // unit-weight kernels never read a weight: nothing to build or upload
    std::vector<int32_t> nbr((size_t)slots), edge((size_t)slots, -1);
    std::vector<double> wv(unitw ? 0 : (size_t)slots, 0.0);
    for (int64_t x = 0; x < V; ++x) {
      int64_t base = (int64_t)soff[x >> 5] * GDS_SLICE + (x & 31);
      int64_t width = soff[(x >> 5) + 1] - soff[x >> 5];
      int64_t deg = h.n_ptr[x + 1] - h.n_ptr[x];
      for (int64_t j = 0; j < width; ++j) {
        int64_t slot = base + j * GDS_SLICE;
        if (j < deg) {
          int64_t a = h.n_ptr[x] + j;
          nbr[slot] = h.n_nbr[a];
          edge[slot] = (h.n_edge[a] << 1) | (h.n_sign[a] > 0 ? 1 : 0);
          if (!unitw) wv[slot] = h.w[h.n_edge[a]];
        } else {
          nbr[slot] = (int32_t)x;
        }
      }
    }
    // lanes of a partial last slice that are not rows: keep gathers in range
    for (int64_t x = V; x < ((V + 31) / 32) * 32; ++x) {
      int64_t base = (int64_t)soff[x >> 5] * GDS_SLICE + (x & 31);
      int64_t width = soff[(x >> 5) + 1] - soff[x >> 5];
      for (int64_t j = 0; j < width; ++j) nbr[base + j * GDS_SLICE] = 0;
    }
    {  // how do the quads (4 consecutive rows) of this graph gather?  sampled over the slices
      const int64_t n_slices = (V + GDS_SLICE - 1) / GDS_SLICE;
      const int64_t step = std::max<int64_t>(1, n_slices / 4096);
      int64_t runs_good = 0, runs_odd = 0, total = 0;
      for (int64_t sl = 0; sl < n_slices; sl += step) {
        const int64_t width = soff[sl + 1] - soff[sl];
        for (int64_t j = 0; j < width; ++j)
          for (int64_t q4 = 0; q4 < GDS_SLICE; q4 += 4) {
            const int64_t row = sl * GDS_SLICE + q4;
            if (row + 3 >= V) continue;
            const int64_t slot = (int64_t)soff[sl] * GDS_SLICE + j * GDS_SLICE + q4;
            const int64_t k0 = nbr[slot];
            ++total;
            if (nbr[slot + 1] != k0 + 1 || nbr[slot + 2] != k0 + 2 || nbr[slot + 3] != k0 + 3) continue;
            const int64_t dlt = k0 - row;
            if ((dlt & 3) == 0 || dlt == 1 || dlt == -1) ++runs_good; else ++runs_odd;
          }
      }
      d.n_quad_odd = (total > 0 && runs_odd * 4 > total && runs_odd > runs_good) ? 1 : 0;
    }
    GDS_TRY(upload(g, soff, &d.n_soff));
    GDS_TRY(upload(g, nbr, &d.n_nbr));
    {  // 16-bit relative neighbour indices.  A difference that does not fit (the ghost rows of a partition sit behind
       // the owned rows: one lattice row per slab; the closing edge of a ring) is stored as the escape value -32768 and
       // the kernel takes that quad's indices from the int32 array; used when fewer than 1 slot in 64 escapes
      const int64_t n_slices = (V + GDS_SLICE - 1) / GDS_SLICE;
      std::vector<int16_t> rel((size_t)slots, 0);
      int64_t escapes = 0;
      for (int64_t sl = 0; sl < n_slices; ++sl) {
        const int64_t width = soff[sl + 1] - soff[sl];
        for (int64_t j = 0; j < width; ++j)
          for (int64_t lane = 0; lane < GDS_SLICE; ++lane) {
            const int64_t row = sl * GDS_SLICE + lane;
            const int64_t slot = (int64_t)soff[sl] * GDS_SLICE + j * GDS_SLICE + lane;
            const int64_t dlt = (row < V) ? (int64_t)nbr[slot] - row : 0;   // lanes past the last row point at themselves
            if (dlt < -32767 || dlt > 32767) { rel[slot] = INT16_MIN; ++escapes; }
            else rel[slot] = (int16_t)dlt;
          }
      }
      const bool use16 = V > 0 && !getenv("GDS_B200_NO_IDX16") && escapes <= slots / 64;
      if (use16) GDS_TRY(upload(g, rel, &d.n_nbr16));
    }
    GDS_TRY(upload(g, edge, &d.n_edge));
    if (!unitw) GDS_TRY(upload(g, wv, &d.n_w));
    GDS_TRY(upload(g, h.n_dinv, &d.n_dinv));
  }
  GDS_TRY(upload(g, h.eu, &d.e_u));
  GDS_TRY(upload(g, h.ev, &d.e_v));
  if (!h.unit_weights) {
    GDS_TRY(upload(g, h.omega, &d.e_omega));
    GDS_TRY(upload(g, h.w, &d.e_w));
  }
  // edge -> faces and face -> edges
  auto pack = [](int64_t n_rows, const std::vector<int64_t>& ptr, const std::vector<int32_t>& idx,
                 const std::vector<int8_t>& sign, std::vector<uint32_t>& soff, std::vector<int32_t>& packed) {
    sell_offsets(n_rows, ptr, soff);
    int64_t slots = (int64_t)soff[soff.size() - 1] * GDS_SLICE;
    packed.assign((size_t)slots, -1);
    for (int64_t r = 0; r < n_rows; ++r) {
      int64_t base = (int64_t)soff[r >> 5] * GDS_SLICE + (r & 31);
      for (int64_t a = ptr[r]; a < ptr[r + 1]; ++a)
        packed[base + (a - ptr[r]) * GDS_SLICE] = (idx[a] << 1) | (sign[a] < 0 ? 1 : 0);
    }
  };
  {
    std::vector<uint32_t> soff;
    std::vector<int32_t> packed;
    pack(E, h.ef_ptr, h.ef_face, h.ef_sign, soff, packed);
    for (size_t i = 0; i + 1 < soff.size(); ++i) d.ef_maxw = std::max<int32_t>(d.ef_maxw, (int32_t)(soff[i + 1] - soff[i]));
    GDS_TRY(upload(g, soff, &d.ef_soff));
    GDS_TRY(upload(g, packed, &d.ef_face));
    pack(F, h.f_ptr, h.f_edge, h.f_sign, soff, packed);
    GDS_TRY(upload(g, soff, &d.fe_soff));
    GDS_TRY(upload(g, packed, &d.fe_edge));
  }
  *out = g;
  return 0;
}

extern "C" int gds_graph_destroy(gds_graph* g) {
  if (!g) return 0;
  if (g->ctx->device < 0) { delete g; return 0; }
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  for (void* p : g->allocs) cudaFree(p);
  delete g;
  return 0;
}

extern "C" int gds_graph_sizes(gds_graph* g, int64_t* V, int64_t* E, int64_t* F) {
  if (!g) GDS_FAIL("null graph");
  if (V) *V = g->h.V;
  if (E) *E = g->h.E;
  if (F) *F = g->h.F;
  return 0;
}

extern "C" int gds_graph_device_bytes(gds_graph* g, int64_t* bytes) {
  if (!g) GDS_FAIL("null graph");
  if (bytes) *bytes = g->device_bytes;
  return 0;
}

extern "C" int gds_graph_export_csr(gds_graph* g, int32_t kind, int64_t* n_rows, int64_t* nnz, int64_t* ptr,
                                    int32_t* idx, int32_t* sign) {
  if (!g) GDS_FAIL("null graph");
  const HostGraph& h = g->h;
  const std::vector<int64_t>* p;
  const std::vector<int32_t>* i;
  const std::vector<int8_t>* s;
  int64_t rows;
  if (kind == 0) { p = &h.n_ptr; i = &h.n_edge; s = &h.n_sign; rows = h.V; }
  else if (kind == 1) { p = &h.f_ptr; i = &h.f_edge; s = &h.f_sign; rows = h.F; }
  else if (kind == 2) { p = &h.ef_ptr; i = &h.ef_face; s = &h.ef_sign; rows = h.E; }
  else GDS_FAIL("unknown kind");
  if (n_rows) *n_rows = rows;
  if (nnz) *nnz = (int64_t)i->size();
  if (ptr) std::copy(p->begin(), p->end(), ptr);
  if (idx) std::copy(i->begin(), i->end(), idx);
  if (sign) for (size_t a = 0; a < s->size(); ++a) sign[a] = (*s)[a];
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// pass resolution
// ---------------------------------------------------------------------------------------------------
static int64_t domain_rows(const gds_graph* g, int32_t domain) {
  if (domain == GDS_NODES) return g->h.V;
  if (domain == GDS_EDGES) return g->h.E;
  if (domain == GDS_FACES) return g->h.F;
  return -1;
}

static int copy_pass(const gds_pass_desc* p, PassHost& ph) {
  if (!p) GDS_FAIL("null pass");
  if (p->n_code < 0 || p->n_code > GDS_MAX_CODE) GDS_FAIL("row program too long");
  if (p->n_imm < 0 || p->n_imm > GDS_MAX_IMM) GDS_FAIL("too many immediates");
  if (p->n_vec < 0 || p->n_vec > GDS_MAX_VEC) GDS_FAIL("too many vector operands");
  if (p->n_mask < 0 || p->n_mask > GDS_MAX_MASK) GDS_FAIL("too many masks");
  if (p->n_out < 0 || p->n_out > GDS_MAX_OUT) GDS_FAIL("too many outputs");
  ph.domain = p->domain; ph.when = p->when;
  ph.code.assign(p->code, p->code + p->n_code);
  ph.imm.assign(p->imm, p->imm + p->n_imm);
  ph.vec.assign(p->vec, p->vec + p->n_vec);
  ph.mask.assign(p->mask, p->mask + p->n_mask);
  ph.out.assign(p->out, p->out + p->n_out);
  ph.out_offset.assign(p->n_out, 0);
  if (p->out_offset) ph.out_offset.assign(p->out_offset, p->out_offset + p->n_out);
  ph.epilogue = p->epilogue;
  ph.state_offset = p->state_offset;
  ph.dirichlet = p->dirichlet;
  // static validation of the row program: opcodes, operand indices, stack depth
  int depth = 0, maxdepth = 0;
  for (uint32_t ins : ph.code) {
    uint32_t op = ins & 0xff, a = (ins >> 8) & 0xff, b = (ins >> 16) & 0xff;
    int push = 0, pop = 0;
    switch (op) {
      case GDS_OP_END: break;
      case GDS_OP_PUSH_VEC: if ((int)a >= p->n_vec) GDS_FAIL("vector operand index out of range"); push = 1; break;
      case GDS_OP_PUSH_IMM: if ((int)a >= p->n_imm) GDS_FAIL("immediate index out of range"); push = 1; break;
      case GDS_OP_PUSH_T: case GDS_OP_PUSH_H: push = 1; break;
      case GDS_OP_DUP: pop = 1; push = 2; break;
      case GDS_OP_SWAP: pop = 2; push = 2; break;
      case GDS_OP_POP: pop = 1; break;
      case GDS_OP_SETL: if (a >= 4) GDS_FAIL("local index out of range"); pop = 1; push = 1; break;
      case GDS_OP_GETL: if (a >= 4) GDS_FAIL("local index out of range"); push = 1; break;
      case GDS_OP_STORE: if ((int)a >= p->n_out) GDS_FAIL("output index out of range"); pop = 1; push = 1; break;
      case GDS_OP_MASK_ZERO: if ((int)a >= p->n_mask) GDS_FAIL("mask index out of range"); pop = 1; push = 1; break;
      case GDS_OP_ADD: case GDS_OP_SUB: case GDS_OP_MUL: case GDS_OP_DIV: case GDS_OP_POW: case GDS_OP_MAX:
      case GDS_OP_MIN: case GDS_OP_ATAN2: pop = 2; push = 1; break;
      case GDS_OP_NEG: case GDS_OP_ABS: case GDS_OP_SIGN: case GDS_OP_SQRT: case GDS_OP_EXP: case GDS_OP_LOG:
      case GDS_OP_SIN: case GDS_OP_COS: case GDS_OP_TAN: case GDS_OP_TANH: case GDS_OP_SINH: case GDS_OP_COSH:
      case GDS_OP_ATAN: case GDS_OP_SQUARE: case GDS_OP_RECIP: case GDS_OP_POWI: pop = 1; push = 1; break;
      case GDS_OP_N_LAP: case GDS_OP_N_BDOT: case GDS_OP_N_INFLUX: case GDS_OP_N_OUTFLUX:
        if (p->domain != GDS_NODES) GDS_FAIL("node gather in a non-node pass");
        if ((int)a >= p->n_vec) GDS_FAIL("vector operand index out of range");
        if (op == GDS_OP_N_LAP && (ins >> 24)) {   // pointwise chain on the gathered values
          int n = (int)(ins >> 24);
          if (n > GDS_MAX_CHAIN) GDS_FAIL("pointwise chain longer than GDS_MAX_CHAIN");
          if ((int)b + 2 * n > p->n_imm) GDS_FAIL("pointwise chain outside the immediates");
          for (int k = 0; k < n; ++k) {
            double code = p->imm[b + 2 * k];
            if (!(code >= GDS_CHAIN_SQUARE && code <= GDS_CHAIN_CUBE) || code != (double)(int)code)
              GDS_FAIL("unknown pointwise chain step");
          }
        }
        push = 1; break;
      case GDS_OP_N_ADVECT: case GDS_OP_N_LIE:
        if (p->domain != GDS_NODES) GDS_FAIL("node gather in a non-node pass");
        if ((int)a >= p->n_vec || (int)b >= p->n_vec) GDS_FAIL("vector operand index out of range");
        push = 1; break;
      case GDS_OP_N_EAGG:
        if (p->domain != GDS_NODES) GDS_FAIL("node gather in a non-node pass");
        if ((int)a >= p->n_vec || (int)b >= p->n_out) GDS_FAIL("operand index out of range");
        break;
      case GDS_OP_N_EAGGT:
        if (p->domain != GDS_NODES) GDS_FAIL("node gather in a non-node pass");
        if ((int)a >= p->n_vec || (int)b >= p->n_out || (int)(ins >> 24) >= p->n_vec) GDS_FAIL("operand index out of range");
        break;
      case GDS_OP_E_GRADT: case GDS_OP_E_CURLT:
        if (p->domain != GDS_EDGES) GDS_FAIL("edge gather in a non-edge pass");
        if ((int)a >= p->n_vec) GDS_FAIL("vector operand index out of range");
        push = 1; break;
      case GDS_OP_E_ADVFIN: case GDS_OP_E_ADVS1:
        if (p->domain != GDS_EDGES) GDS_FAIL("edge gather in a non-edge pass");
        if ((int)a >= p->n_vec || (int)b >= p->n_vec) GDS_FAIL("vector operand index out of range");
        push = 1; break;
      case GDS_OP_F_CURL:
        if (p->domain != GDS_FACES) GDS_FAIL("face gather in a non-face pass");
        if ((int)a >= p->n_vec) GDS_FAIL("vector operand index out of range");
        push = 1; break;
      default: GDS_FAIL("unknown opcode " + std::to_string(op));
    }
    if (depth < pop) GDS_FAIL("row program pops an empty stack");
    depth += push - pop;
    maxdepth = std::max(maxdepth, depth);
  }
  if (maxdepth > 6) GDS_FAIL("row program needs a stack deeper than 6");
  if (ph.epilogue == GDS_EPI_RHS && depth < 1) GDS_FAIL("RHS pass leaves nothing on the stack");
  return 0;
}

// Recognise row programs that are a linear combination of masked one-hop operators,
//   value = sum_k coef_k * Z(op_k(...)) + beta * b + gamma,
// by running the program on symbolic forms.  Anything outside that family (time, nonlinear pointwise operations,
// products of vectors) is left to the interpreter.  A single N_LAP term is the LapForm of the Laplacian kernels.
static bool is_gather(uint32_t op) {
  switch (op) {
    case GDS_OP_N_LAP: case GDS_OP_N_BDOT: case GDS_OP_N_INFLUX: case GDS_OP_N_OUTFLUX: case GDS_OP_N_ADVECT:
    case GDS_OP_N_LIE: case GDS_OP_E_GRADT: case GDS_OP_E_CURLT: case GDS_OP_E_ADVFIN: case GDS_OP_F_CURL: return true;
    default: return false;
  }
}

static void match_forms(PassHost& ph) {
  ph.lap = LapForm();
  ph.terms = TermForm();
  struct Sym {
    std::vector<TermSym> t;
    double beta = 0, beta2 = 0, gamma = 0;
    int b = -1, b2 = -1;     // up to two additive vectors; only the first may carry a pointwise chain
    Chain bchain = {};
  };
  auto same_chain = [](const Chain& x, const Chain& y) {
    if (x.n != y.n) return false;
    for (int i = 0; i < x.n; ++i) if (x.pre[i] != y.pre[i] || x.m[i] != y.m[i] || x.a[i] != y.a[i]) return false;
    return true;
  };
  std::vector<Sym> st;
  int store = -1;
  auto scalar = [](const Sym& x) { return x.t.empty() && x.b < 0; };
  auto scale = [](Sym& x, double c) {
    for (auto& t : x.t) t.coef *= c;
    x.beta *= c; x.beta2 *= c; x.gamma *= c;
  };
  // x += sg * (beta * g(b))  for one additive vector of y; false if x has no room for it
  auto add_vec = [&](Sym& x, int slot, double beta, const Chain& ch) -> bool {
    if (slot < 0) return true;
    if (x.b == slot && same_chain(x.bchain, ch)) { x.beta += beta; return true; }
    if (x.b2 == slot && ch.n == 0) { x.beta2 += beta; return true; }
    if (x.b < 0) { x.b = slot; x.beta = beta; x.bchain = ch; return true; }
    if (x.b2 < 0 && ch.n == 0) { x.b2 = slot; x.beta2 = beta; return true; }
    if (x.b2 < 0 && x.bchain.n == 0) {   // keep the chained vector in the first place
      x.b2 = x.b; x.beta2 = x.beta; x.b = slot; x.beta = beta; x.bchain = ch; return true;
    }
    return false;
  };
  for (size_t pc = 0; pc < ph.code.size(); ++pc) {
    uint32_t ins = ph.code[pc], op = ins & 0xff, a = (ins >> 8) & 0xff, b = (ins >> 16) & 0xff;
    if (is_gather(op)) {
      Sym x;
      TermSym t;
      t.op = (int)op; t.a = (int)a; t.coef = 1.0;
      if (op == GDS_OP_N_ADVECT || op == GDS_OP_N_LIE || op == GDS_OP_E_ADVFIN) t.b = (int)b;
      if (op == GDS_OP_N_LAP) {
        for (int k = 0; k < (int)(ins >> 24); ++k)
          if (!chain_push(t.chain, (int)ph.imm[b + 2 * k], ph.imm[b + 2 * k + 1])) return;
      }
      x.t.push_back(t);
      st.push_back(x);
      continue;
    }
    switch (op) {
      case GDS_OP_END: break;
      case GDS_OP_PUSH_IMM: { Sym x; x.gamma = ph.imm[a]; st.push_back(x); } break;
      case GDS_OP_PUSH_VEC: { Sym x; x.beta = 1.0; x.b = (int)a; st.push_back(x); } break;
      case GDS_OP_MASK_ZERO: {
        if (st.empty()) return;
        Sym& x = st.back();
        if (x.t.empty() || x.b >= 0 || x.gamma != 0.0) return;
        for (auto& t : x.t) {
          if (t.m0 == (int)a || t.m1 == (int)a) continue;
          if (t.m0 < 0) t.m0 = (int)a;
          else if (t.m1 < 0) t.m1 = (int)a;
          else return;
        }
      } break;
      case GDS_OP_NEG: {
        if (st.empty()) return;
        scale(st.back(), -1.0);
      } break;
      case GDS_OP_SWAP: {
        if (st.size() < 2) return;
        std::swap(st[st.size() - 1], st[st.size() - 2]);
      } break;
      case GDS_OP_POWI: {
        // b[row]**2 / **3 of a bare vector: the same pointwise chain step as SQUARE / the interpreter's powi
        if (st.empty()) return;
        Sym& x = st.back();
        const int n = (int)(int8_t)a;
        if ((n != 2 && n != 3) || !x.t.empty() || x.b < 0 || x.b2 >= 0 || x.beta != 1.0 || x.gamma != 0.0) return;
        if (!chain_push(x.bchain, n == 2 ? GDS_CHAIN_SQUARE : GDS_CHAIN_CUBE, 0.0)) return;
      } break;
      case GDS_OP_SQUARE: case GDS_OP_ABS: {
        // g(b[row]) for a bare vector: a pointwise chain on the additive operand (beta * g(b))
        if (st.empty()) return;
        Sym& x = st.back();
        if (!x.t.empty() || x.b < 0 || x.b2 >= 0 || x.beta != 1.0 || x.gamma != 0.0) return;
        if (!chain_push(x.bchain, (op == GDS_OP_SQUARE) ? GDS_CHAIN_SQUARE : GDS_CHAIN_ABS, 0.0)) return;
      } break;
      case GDS_OP_ADD: case GDS_OP_SUB: {
        if (st.size() < 2) return;
        Sym y = st.back(); st.pop_back();
        Sym& x = st.back();
        double sg = (op == GDS_OP_SUB) ? -1.0 : 1.0;
        for (auto& ty : y.t) {
          bool merged = false;
          for (auto& tx : x.t)
            if (tx.op == ty.op && tx.a == ty.a && tx.b == ty.b && tx.m0 == ty.m0 && tx.m1 == ty.m1 &&
                same_chain(tx.chain, ty.chain)) {
              tx.coef += sg * ty.coef; merged = true; break;
            }
          if (!merged) { TermSym t = ty; t.coef = sg * ty.coef; x.t.push_back(t); }
        }
        if (!add_vec(x, y.b, sg * y.beta, y.bchain)) return;
        const Chain none = {};
        if (!add_vec(x, y.b2, sg * y.beta2, none)) return;
        x.gamma += sg * y.gamma;
      } break;
      case GDS_OP_MUL: {
        if (st.size() < 2) return;
        Sym y = st.back(); st.pop_back();
        Sym& x = st.back();
        if (scalar(y)) scale(x, y.gamma);
        else if (scalar(x)) { double c = x.gamma; x = y; scale(x, c); }
        else return;
      } break;
      case GDS_OP_DIV: {
        if (st.size() < 2) return;
        Sym y = st.back(); st.pop_back();
        if (!scalar(y)) return;
        Sym& x = st.back();
        for (auto& t : x.t) t.coef /= y.gamma;
        x.beta /= y.gamma; x.beta2 /= y.gamma; x.gamma /= y.gamma;
      } break;
      case GDS_OP_STORE:
        if (pc + 1 != ph.code.size() || store >= 0) return;   // only as the last instruction
        store = (int)a;
        break;
      default: return;
    }
  }
  if (st.size() != 1) return;
  const Sym& r = st.back();
  if ((int)r.t.size() > GDS_MAX_TERMS) return;   // no term at all: a pointwise pass  beta * b + gamma
  if (!std::isfinite(r.beta) || !std::isfinite(r.beta2) || !std::isfinite(r.gamma)) return;
  for (auto& t : r.t) if (!std::isfinite(t.coef)) return;
  const int bslot = (r.b >= 0 && r.beta != 0.0) ? r.b : -1;
  const int b2slot = (r.b2 >= 0 && r.beta2 != 0.0) ? r.b2 : -1;
  if (r.t.size() == 1 && r.t[0].op == GDS_OP_N_LAP && r.t[0].m1 < 0 && r.t[0].coef != 0.0) {
    ph.lap.ok = true;
    ph.lap.v = r.t[0].a; ph.lap.mask = r.t[0].m0; ph.lap.b = bslot;
    ph.lap.alpha = r.t[0].coef; ph.lap.beta = r.beta; ph.lap.gamma = r.gamma; ph.lap.store = store;
    ph.lap.chain = r.t[0].chain;
    if (bslot >= 0) ph.lap.bchain = r.bchain;
    ph.lap.b2 = b2slot; ph.lap.beta2 = r.beta2;
    return;
  }
  ph.terms.ok = true;
  ph.terms.n = (int)r.t.size();
  for (size_t k = 0; k < r.t.size(); ++k) ph.terms.t[k] = r.t[k];
  ph.terms.b = bslot; ph.terms.beta = r.beta; ph.terms.gamma = r.gamma; ph.terms.store = store;
  if (bslot >= 0) ph.terms.bchain = r.bchain;
  ph.terms.b2 = b2slot; ph.terms.beta2 = r.beta2;
}

// Resolve a host pass into launch parameters.  `state`/`stage` may be null for eager passes.
static int resolve_pass(const gds_graph* g, const PassHost& ph, const double* state, const double* stage,
                        PassParams& pp) {
  memset(&pp, 0, sizeof(pp));
  pp.g = g->d;
  pp.n_rows = domain_rows(g, ph.domain);
  if (pp.n_rows < 0) GDS_FAIL("unknown domain");
  pp.n_code = (int32_t)ph.code.size();
  for (size_t i = 0; i < ph.code.size(); ++i) pp.code[i] = ph.code[i];
  for (size_t i = 0; i < ph.imm.size(); ++i) pp.imm[i] = ph.imm[i];
  for (size_t i = 0; i < ph.vec.size(); ++i) {
    const gds_vecref& v = ph.vec[i];
    if (v.kind == GDS_VEC_BUF) {
      if (!v.buf) GDS_FAIL("null buffer operand");
      if (v.offset < 0 || v.offset > v.buf->n) GDS_FAIL("operand offset outside buffer");
      pp.vec[i] = v.buf->d + v.offset;
    } else if (v.kind == GDS_VEC_STATE) {
      if (!state) GDS_FAIL("state operand outside a system");
      pp.vec[i] = state + v.offset;
    } else if (v.kind == GDS_VEC_STAGE) {
      if (!stage) GDS_FAIL("stage operand outside a system");
      pp.vec[i] = stage + v.offset;
    } else {
      GDS_FAIL("unknown operand kind");
    }
  }
  for (size_t i = 0; i < ph.mask.size(); ++i) {
    if (!ph.mask[i]) GDS_FAIL("null mask");
    if (ph.mask[i]->n_rows < pp.n_rows) GDS_FAIL("mask shorter than the pass domain");
    pp.mask[i] = ph.mask[i]->d;
  }
  for (size_t i = 0; i < ph.out.size(); ++i) {
    if (!ph.out[i]) GDS_FAIL("null output buffer");
    pp.out[i] = ph.out[i]->d + ph.out_offset[i];
  }
  pp.epilogue = ph.epilogue;
  pp.unit_weights = g->h.unit_weights ? 1 : 0;
  if (ph.lap.ok && !ph.special) {
    pp.lap_v = pp.vec[ph.lap.v];
    pp.lap_mask = ph.lap.mask >= 0 ? pp.mask[ph.lap.mask] : nullptr;
    pp.lap_b = ph.lap.b >= 0 ? pp.vec[ph.lap.b] : nullptr;
    pp.lap_store = ph.lap.store >= 0 ? pp.out[ph.lap.store] : nullptr;
    pp.lap_alpha = ph.lap.alpha; pp.lap_beta = ph.lap.beta; pp.lap_gamma = ph.lap.gamma;
    pp.lap_chain = ph.lap.chain; pp.lap_bchain = ph.lap.bchain;
    pp.lap_b2 = ph.lap.b2 >= 0 ? pp.vec[ph.lap.b2] : nullptr;
    pp.lap_beta2 = ph.lap.beta2;
  } else if (ph.terms.ok) {
    pp.use_terms = 1;
    pp.n_terms = ph.terms.n;
    for (int k = 0; k < ph.terms.n; ++k) {
      const TermSym& t = ph.terms.t[k];
      pp.term[k].op = t.op;
      pp.term[k].a = pp.vec[t.a];
      pp.term[k].b = t.b >= 0 ? pp.vec[t.b] : nullptr;
      pp.term[k].m0 = t.m0 >= 0 ? pp.mask[t.m0] : nullptr;
      pp.term[k].m1 = t.m1 >= 0 ? pp.mask[t.m1] : nullptr;
      pp.term[k].coef = t.coef;
      pp.term[k].chain = t.chain;
    }
    pp.lap_bchain = ph.terms.bchain;
    pp.lap_b2 = ph.terms.b2 >= 0 ? pp.vec[ph.terms.b2] : nullptr;
    pp.lap_beta2 = ph.terms.beta2;
    pp.lap_b = ph.terms.b >= 0 ? pp.vec[ph.terms.b] : nullptr;
    pp.lap_store = ph.terms.store >= 0 ? pp.out[ph.terms.store] : nullptr;
    pp.lap_beta = ph.terms.beta; pp.lap_gamma = ph.terms.gamma;
  }
  return 0;
}

// which kernel family a pass runs in: 0 row-program interpreter, 1 Laplacian kernels (lapq / lapT), 2 terms_kernel.
// Pure host logic (works on a host-only context): lets the parity tests pin the dispatch of every BASELINE law.
extern "C" int gds_pass_kernel_kind(const gds_pass_desc* pass, int32_t* kind_out) {
  if (!kind_out) GDS_FAIL("null out");
  PassHost ph;
  GDS_TRY(copy_pass(pass, ph));
  match_forms(ph);
  *kind_out = ph.lap.ok ? 1 : (ph.terms.ok ? 2 : 0);
  return 0;
}

extern "C" int gds_pass_run(gds_ctx* c, gds_graph* g, const gds_pass_desc* pass, double t) {
  if (!c || !g) GDS_FAIL("bad argument");
  PassHost ph;
  GDS_TRY(copy_pass(pass, ph));   // validation works on a host-only context too
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(ph);
  if (ph.epilogue != GDS_EPI_NONE) GDS_FAIL("eager passes take GDS_EPI_NONE");
  PassParams pp;
  GDS_TRY(resolve_pass(g, ph, nullptr, nullptr, pp));
  pp.t_fixed = t;
  return launch_pass(c, pp, ph.domain);
}

// Matrix operands (operators applied to `np.eye(n)`, examples/fluid.py:350; blocks of tangent vectors for advect_jvp): the
// pass list is validated, matched and resolved ONCE, then enqueued for every column with the operands that hold one
// column per `stride` advanced accordingly -- one call, one upload and one download for the whole matrix.
extern "C" int gds_passes_run_cols(gds_ctx* c, gds_graph* g, int32_t n_pass, const gds_pass_desc* passes, double t,
                                   int64_t ncols, int32_t n_strided, gds_buf* const* strided, const int64_t* stride) {
  if (!c || !g || n_pass < 0 || (n_pass > 0 && !passes) || ncols < 0 || n_strided < 0) GDS_FAIL("bad argument");
  std::vector<PassHost> ph((size_t)n_pass);
  for (int i = 0; i < n_pass; ++i) {
    GDS_TRY(copy_pass(&passes[i], ph[(size_t)i]));
    if (ph[(size_t)i].epilogue != GDS_EPI_NONE) GDS_FAIL("eager passes take GDS_EPI_NONE");
  }
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  auto stride_of = [&](const gds_buf* b) -> int64_t {
    for (int k = 0; k < n_strided; ++k) if (strided[k] == b) return stride[k];
    return 0;
  };
  for (int k = 0; k < n_strided; ++k)
    if (!strided[k] || stride[k] < 0 || stride[k] * ncols > strided[k]->n) GDS_FAIL("column stride outside its buffer");
  for (auto& p : ph) if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(p);
  for (int64_t col = 0; col < ncols; ++col) {
    for (auto& p : ph) {
      PassHost q = p;     // operands of this column
      for (auto& v : q.vec) if (v.kind == GDS_VEC_BUF && v.buf) v.offset += col * stride_of(v.buf);
      for (size_t j = 0; j < q.out.size(); ++j) q.out_offset[j] += col * stride_of(q.out[j]);
      PassParams pp;
      GDS_TRY(resolve_pass(g, q, nullptr, nullptr, pp));
      pp.t_fixed = t;
      GDS_TRY(launch_pass(c, pp, q.domain));
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// conjugate gradients with the operator given as a pass
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_cg_solve(gds_ctx* c, gds_graph* g, const gds_pass_desc* apply, gds_buf* x, gds_buf* b, gds_buf* r,
                            gds_buf* p, gds_buf* ap, int64_t n, double rtol, int32_t max_iter, int32_t check_every,
                            int32_t* iters_out, double* relres_out) {
  if (!c || !g || !apply || !x || !b || !r || !p || !ap) GDS_FAIL("bad argument");
  if (n < 0 || x->n < n || b->n < n || r->n < n || p->n < n || ap->n < n) GDS_FAIL("vector shorter than n");
  if (max_iter < 0 || check_every < 1 || !(rtol >= 0.0)) GDS_FAIL("bad iteration parameters");
  PassHost ph;
  GDS_TRY(copy_pass(apply, ph));
  if (ph.epilogue != GDS_EPI_NONE) GDS_FAIL("the operator pass takes GDS_EPI_NONE");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(ph);
  PassParams pp;
  GDS_TRY(resolve_pass(g, ph, nullptr, nullptr, pp));
  double* S = nullptr;
  double* partial = nullptr;
  GDS_CUDA(cudaMalloc((void**)&S, 16 * sizeof(double)));
  if (cudaMalloc((void**)&partial, 4096 * sizeof(double)) != cudaSuccess) { cudaFree(S); GDS_FAIL("out of device memory"); }
  int rc = 0, it = 0;
  double h[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool fused = !getenv("GDS_B200_CG_UNFUSED");   // 4 launches per iteration instead of 6 (same arithmetic, same bits)
  auto fetch = [&]() -> int {
    GDS_CUDA(cudaMemcpyAsync(h, S, 10 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GDS_CUDA(cudaStreamSynchronize(c->stream));
    if (fused && it > 0) {   // the last direction update left its scalars in the shadow slots
      h[0] = h[7];
      if (h[8] != 0.0 || h[9] != 0.0) h[6] = 1.0;
    }
    return 0;
  };
  // x = 0, r = b, p = b, rr = bb = b.b
  rc = cudaMemsetAsync(x->d, 0, (size_t)n * 8, c->stream) != cudaSuccess;
  rc = rc || cudaMemcpyAsync(r->d, b->d, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess;
  rc = rc || cudaMemcpyAsync(p->d, b->d, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess;
  rc = rc || cudaMemsetAsync(S, 0, 16 * sizeof(double), c->stream) != cudaSuccess;
  if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, r->d, r->d, n, S, partial, 0);
  if (!rc) rc = launch_cg(c, 1, nullptr, nullptr, nullptr, nullptr, n, S, partial, 2);
  if (!rc) rc = fetch();
  const double bb = h[4], target = rtol * rtol * bb;
  double best = bb;
  if (!rc) rc = cudaMemcpyAsync(S + 5, &target, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
  if (!rc) rc = cudaMemcpyAsync(S + 7, &bb, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;   // shadow of r.r
  if (!rc) rc = cudaStreamSynchronize(c->stream) != cudaSuccess;
  while (!rc && it < max_iter && h[0] > target && bb > 0.0 && h[6] == 0.0) {
    int m = std::min<int>(check_every, max_iter - it);
    for (int k = 0; k < m && !rc && fused; ++k) {
      rc = launch_pass(c, pp, ph.domain);                                                   // ap = A p
      if (!rc) rc = launch_cg(c, 4, nullptr, nullptr, p->d, ap->d, n, S, partial, 0);       // p.Ap partial sums (+ commit)
      if (!rc) rc = launch_cg(c, 5, x->d, r->d, p->d, ap->d, n, S, partial, 0);             // alpha; x, r; r.r partial sums
      if (!rc) rc = launch_cg(c, 6, nullptr, r->d, p->d, nullptr, n, S, partial, 0);        // beta; p = r + beta p
    }
    for (int k = 0; k < m && !rc && !fused; ++k) {
      rc = launch_pass(c, pp, ph.domain);                                                   // ap = A p
      if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, p->d, ap->d, n, S, partial, 0);       // p.Ap
      if (!rc) rc = launch_cg(c, 1, nullptr, nullptr, nullptr, nullptr, n, S, partial, 0);  // alpha
      if (!rc) rc = launch_cg(c, 2, x->d, r->d, p->d, ap->d, n, S, partial, 0);             // x, r, partial r.r
      if (!rc) rc = launch_cg(c, 1, nullptr, nullptr, nullptr, nullptr, n, S, partial, 1);  // beta, rr
      if (!rc) rc = launch_cg(c, 3, nullptr, r->d, p->d, nullptr, n, S, partial, 0);        // p = r + beta p
    }
    it += m;
    if (!rc) rc = fetch();
    // a right-hand side with a component outside the range of a singular operator stalls and then diverges: stop
    // once the residual has grown well past its best value
    if (!rc) {
      if (!(h[0] == h[0]) || h[0] > 1e4 * best) break;
      best = std::min(best, h[0]);
    }
  }
  cudaFree(S);
  cudaFree(partial);
  if (rc) { if (rc == 1) GDS_FAIL("CUDA failure inside the CG loop"); return rc; }
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = (bb > 0.0) ? std::sqrt(h[0] / bb) : 0.0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// multigrid preconditioner (include/gds_b200.h: gds_mg_*, gds_pcg_solve)
// ---------------------------------------------------------------------------------------------------
struct MgCsr {
  int64_t rows = 0, nnz = 0;
  int32_t* rowptr = nullptr;
  int32_t* col = nullptr;
  double* val = nullptr;
  int tpr = 1;          // threads that share a row
};
struct MgLevel {
  int64_t n = 0, n_coarse = 0;
  MgCsr A, P, R;        // coarsest level: A holds the dense pseudo-inverse, P and R are empty
  double* w = nullptr;  // omega / diag(A)
  double* xa = nullptr; // iterate (ping)
  double* xb = nullptr; // iterate (pong); the level's result when it is not written straight into the caller's buffer
  double* b = nullptr;  // right-hand side (levels > 0: the restricted residual)
  double* r = nullptr;  // residual
};
struct gds_mg {
  gds_ctx* ctx = nullptr;
  std::vector<MgLevel> lev;
  int pre = 1, post = 1;
  int64_t device_bytes = 0;
  bool closed = false;  // the coarsest level has been added
};

static void mg_free_csr(MgCsr& m) { cudaFree(m.rowptr); cudaFree(m.col); cudaFree(m.val); m = MgCsr(); }

static int mg_upload_csr(gds_mg* mg, int64_t rows, int64_t cols, const int64_t* rowptr, const int32_t* col, const double* val,
                         MgCsr& out) {
  if (!rowptr || rowptr[0] != 0) GDS_FAIL("CSR row pointers start at 0");
  const int64_t nnz = rowptr[rows];
  if (nnz < 0 || nnz >= (int64_t(1) << 31) - 64) GDS_FAIL("CSR matrix too large for 32-bit positions");
  if (nnz > 0 && (!col || !val)) GDS_FAIL("CSR columns / values missing");
  std::vector<int32_t> rp((size_t)rows + 1);
  for (int64_t i = 0; i <= rows; ++i) {
    if (i > 0 && rowptr[i] < rowptr[i - 1]) GDS_FAIL("CSR row pointers must not decrease");
    rp[(size_t)i] = (int32_t)rowptr[i];
  }
  for (int64_t k = 0; k < nnz; ++k)
    if (col[k] < 0 || col[k] >= cols) GDS_FAIL("CSR column index out of range");
  out.rows = rows; out.nnz = nnz;
  const double avg = rows > 0 ? (double)nnz / (double)rows : 0.0;
  int tpr = 1;
  while (tpr < 32 && (double)tpr * 2.0 < avg) tpr *= 2;     // about half the mean row length, a power of two
  out.tpr = tpr;
  GDS_CUDA(cudaMalloc((void**)&out.rowptr, (size_t)(rows + 1) * sizeof(int32_t)));
  GDS_CUDA(cudaMalloc((void**)&out.col, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
  GDS_CUDA(cudaMalloc((void**)&out.val, (size_t)std::max<int64_t>(nnz, 1) * sizeof(double)));
  GDS_CUDA(cudaMemcpy(out.rowptr, rp.data(), (size_t)(rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (nnz > 0) {
    GDS_CUDA(cudaMemcpy(out.col, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    GDS_CUDA(cudaMemcpy(out.val, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
  }
  mg->device_bytes += (rows + 1) * 4 + nnz * 12;
  return 0;
}

extern "C" int gds_mg_create(gds_ctx* c, gds_mg** out) {
  if (!c || !out) GDS_FAIL("bad argument");
  GDS_NEED_DEVICE(c);
  gds_mg* mg = new gds_mg();
  mg->ctx = c;
  *out = mg;
  return 0;
}

extern "C" int gds_mg_destroy(gds_mg* mg) {
  if (!mg) return 0;
  cudaSetDevice(mg->ctx->device);
  cudaStreamSynchronize(mg->ctx->stream);
  for (auto& l : mg->lev) {
    mg_free_csr(l.A); mg_free_csr(l.P); mg_free_csr(l.R);
    cudaFree(l.w); cudaFree(l.xa); cudaFree(l.xb); cudaFree(l.b); cudaFree(l.r);
  }
  delete mg;
  return 0;
}

extern "C" int gds_mg_add_level(gds_mg* mg, int64_t n, const int64_t* a_rowptr, const int32_t* a_col, const double* a_val,
                                const double* jacobi_w, int64_t n_coarse, const int64_t* p_rowptr, const int32_t* p_col,
                                const double* p_val, const int64_t* r_rowptr, const int32_t* r_col, const double* r_val) {
  if (!mg || n <= 0 || n_coarse < 0) GDS_FAIL("bad argument");
  if (mg->closed) GDS_FAIL("the coarsest level has already been added");
  if (!mg->lev.empty() && mg->lev.back().n_coarse != n) GDS_FAIL("level size differs from the previous level's coarse size");
  if (n_coarse > 0 && (!jacobi_w || !p_rowptr || !r_rowptr)) GDS_FAIL("a level with a coarser one needs weights, P and R");
  gds_ctx* c = mg->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  mg->lev.emplace_back();
  MgLevel& l = mg->lev.back();
  l.n = n; l.n_coarse = n_coarse;
  GDS_TRY(mg_upload_csr(mg, n, n, a_rowptr, a_col, a_val, l.A));
  const size_t vb = (size_t)(n + 4) * sizeof(double);
  if (n_coarse > 0) {
    GDS_TRY(mg_upload_csr(mg, n, n_coarse, p_rowptr, p_col, p_val, l.P));
    GDS_TRY(mg_upload_csr(mg, n_coarse, n, r_rowptr, r_col, r_val, l.R));
    GDS_CUDA(cudaMalloc((void**)&l.w, vb));
    GDS_CUDA(cudaMemcpy(l.w, jacobi_w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    GDS_CUDA(cudaMalloc((void**)&l.xa, vb));
    GDS_CUDA(cudaMalloc((void**)&l.r, vb));
    mg->device_bytes += 3 * (int64_t)vb;
  } else {
    mg->closed = true;
  }
  GDS_CUDA(cudaMalloc((void**)&l.xb, vb));
  mg->device_bytes += (int64_t)vb;
  if (mg->lev.size() > 1) {       // level 0 takes its right-hand side from the caller
    GDS_CUDA(cudaMalloc((void**)&l.b, vb));
    mg->device_bytes += (int64_t)vb;
  }
  return 0;
}

extern "C" int gds_mg_set_sweeps(gds_mg* mg, int32_t pre, int32_t post) {
  if (!mg || pre < 1 || post < 1 || pre > 8 || post > 8) GDS_FAIL("1 .. 8 sweeps before and after the coarse correction");
  mg->pre = pre; mg->post = post;
  return 0;
}

extern "C" int gds_mg_info(gds_mg* mg, int32_t* n_levels, int64_t* device_bytes, int64_t* launches_per_cycle) {
  if (!mg) GDS_FAIL("null hierarchy");
  if (n_levels) *n_levels = (int32_t)mg->lev.size();
  if (device_bytes) *device_bytes = mg->device_bytes;
  if (launches_per_cycle) *launches_per_cycle = mg->lev.empty() ? 0 : (int64_t)(mg->lev.size() - 1) * (mg->pre + mg->post + 3) + 1;
  return 0;
}

static int mg_spmv(gds_ctx* c, const MgCsr& m, const double* x, const double* u, const double* v, const double* w,
                   double* out, int mode) {
  return launch_mg_spmv(c, m.tpr, m.rows, m.rowptr, m.col, m.val, x, u, v, w, out, mode);
}

// one V-cycle from a zero guess on level `li`: out = M_l^-1 b.  Sweeps alternate between the level's two iterate buffers;
// the LAST post-sweep writes `out`, so nothing is copied.
static int mg_cycle(gds_mg* mg, size_t li, const double* b, double* out) {
  gds_ctx* c = mg->ctx;
  MgLevel& l = mg->lev[li];
  if (l.n_coarse == 0) return mg_spmv(c, l.A, b, nullptr, nullptr, nullptr, out, 0);     // dense pseudo-inverse
  // `out` may be this level's xb (the parent reads the result there): start on the buffer that leaves the iterate in xa
  // before the last sweep, whatever the number of sweeps
  const int flips = (mg->pre - 1) + (mg->post - 1);
  double* cur = (flips % 2 == 0) ? l.xa : l.xb;
  double* other = (cur == l.xa) ? l.xb : l.xa;
  GDS_TRY(launch_mg_scale(c, l.n, l.w, b, cur));                                          // sweep 1 from x = 0
  for (int k = 1; k < mg->pre; ++k) {
    GDS_TRY(mg_spmv(c, l.A, cur, b, cur, l.w, other, 3));
    std::swap(cur, other);
  }
  GDS_TRY(mg_spmv(c, l.A, cur, b, nullptr, nullptr, l.r, 1));                             // r = b - A x
  MgLevel& n = mg->lev[li + 1];
  GDS_TRY(mg_spmv(c, l.R, l.r, nullptr, nullptr, nullptr, n.b, 0));                       // b_c = R r
  GDS_TRY(mg_cycle(mg, li + 1, n.b, n.xb));
  GDS_TRY(mg_spmv(c, l.P, n.xb, cur, nullptr, nullptr, cur, 2));                          // x += P x_c
  for (int k = 0; k < mg->post; ++k) {
    double* dst = (k == mg->post - 1) ? out : other;
    GDS_TRY(mg_spmv(c, l.A, cur, b, cur, l.w, dst, 3));
    other = cur;
    cur = dst;
  }
  return 0;
}

static int mg_check(gds_mg* mg, int64_t n) {
  if (!mg || mg->lev.empty() || !mg->closed) GDS_FAIL("the hierarchy is not complete (add levels down to the coarsest one)");
  if (mg->lev[0].n != n) GDS_FAIL("vector length differs from the finest level");
  return 0;
}

extern "C" int gds_mg_apply(gds_mg* mg, gds_buf* r, gds_buf* z) {
  if (!mg || !r || !z) GDS_FAIL("bad argument");
  GDS_TRY(mg_check(mg, mg->lev.empty() ? -1 : mg->lev[0].n));
  if (r->n < mg->lev[0].n || z->n < mg->lev[0].n || r->d == z->d) GDS_FAIL("need two distinct vectors of the finest level's size");
  GDS_CUDA(cudaSetDevice(mg->ctx->device));
  return mg_cycle(mg, 0, r->d, z->d);
}

extern "C" int gds_pcg_solve(gds_ctx* c, gds_graph* g, const gds_pass_desc* apply, gds_mg* mg, gds_buf* x, gds_buf* b,
                             gds_buf* r, gds_buf* p, gds_buf* ap, gds_buf* z, int64_t n, double rtol, int32_t max_iter,
                             int32_t check_every, int32_t* iters_out, double* relres_out) {
  if (!c || !g || !apply || !mg || !x || !b || !r || !p || !ap || !z) GDS_FAIL("bad argument");
  if (n < 0 || x->n < n || b->n < n || r->n < n || p->n < n || ap->n < n || z->n < n) GDS_FAIL("vector shorter than n");
  if (max_iter < 0 || check_every < 1 || !(rtol >= 0.0)) GDS_FAIL("bad iteration parameters");
  if (mg->ctx != c) GDS_FAIL("the hierarchy belongs to another context");
  GDS_TRY(mg_check(mg, n));
  PassHost ph;
  GDS_TRY(copy_pass(apply, ph));
  if (ph.epilogue != GDS_EPI_NONE) GDS_FAIL("the operator pass takes GDS_EPI_NONE");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(ph);
  PassParams pp;
  GDS_TRY(resolve_pass(g, ph, nullptr, nullptr, pp));
  double* S = nullptr;
  double* partial = nullptr;
  GDS_CUDA(cudaMalloc((void**)&S, 16 * sizeof(double)));
  if (cudaMalloc((void**)&partial, 2048 * sizeof(double)) != cudaSuccess) { cudaFree(S); GDS_FAIL("out of device memory"); }
  int rc = 0, it = 0;
  double h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto fetch = [&]() -> int {
    GDS_CUDA(cudaMemcpyAsync(h, S, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GDS_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };
  // x = 0, r = b, z = M^-1 r, p = z, |r|^2 = |b|^2, r.z
  rc = cudaMemsetAsync(x->d, 0, (size_t)n * 8, c->stream) != cudaSuccess;
  rc = rc || cudaMemcpyAsync(r->d, b->d, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess;
  rc = rc || cudaMemsetAsync(S, 0, 16 * sizeof(double), c->stream) != cudaSuccess;
  if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, r->d, r->d, n, S, partial, 0);
  if (!rc) rc = launch_pcg_final(c, partial, S, 3);
  if (!rc) rc = mg_cycle(mg, 0, r->d, z->d);
  if (!rc) rc = cudaMemcpyAsync(p->d, z->d, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess;
  if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, r->d, z->d, n, S, partial, 0);
  if (!rc) rc = launch_pcg_final(c, partial, S, 4);
  if (!rc) rc = fetch();
  const double bb = h[4], target = rtol * rtol * bb;
  double best = bb;
  if (!rc) rc = cudaMemcpyAsync(S + 5, &target, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
  if (!rc) rc = cudaStreamSynchronize(c->stream) != cudaSuccess;
  while (!rc && it < max_iter && h[7] > target && bb > 0.0 && h[6] == 0.0) {
    int m = std::min<int>(check_every, max_iter - it);
    for (int k = 0; k < m && !rc; ++k) {
      rc = launch_pass(c, pp, ph.domain);                                                     // ap = A p
      if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, p->d, ap->d, n, S, partial, 0);         // p.Ap
      if (!rc) rc = launch_pcg_final(c, partial, S, 0);                                       // alpha = r.z / p.Ap
      if (!rc) rc = launch_cg(c, 2, x->d, r->d, p->d, ap->d, n, S, partial, 0);               // x, r, partial |r|^2
      if (!rc) rc = launch_pcg_final(c, partial, S, 1);                                       // |r|^2, finished?
      if (!rc) rc = mg_cycle(mg, 0, r->d, z->d);                                              // z = M^-1 r
      if (!rc) rc = launch_cg(c, 0, nullptr, nullptr, r->d, z->d, n, S, partial, 0);          // r.z
      if (!rc) rc = launch_pcg_final(c, partial, S, 2);                                       // beta
      if (!rc) rc = launch_cg(c, 3, nullptr, z->d, p->d, nullptr, n, S, partial, 0);          // p = z + beta p
    }
    it += m;
    if (!rc) rc = fetch();
    if (!rc) {
      if (!(h[7] == h[7]) || h[7] > 1e4 * best) break;
      best = std::min(best, h[7]);
    }
  }
  cudaFree(S);
  cudaFree(partial);
  if (rc) { if (rc == 1) GDS_FAIL("CUDA failure inside the preconditioned CG loop"); return rc; }
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = (bb > 0.0) ? std::sqrt(h[7] / bb) : 0.0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// systems
// ---------------------------------------------------------------------------------------------------
#define GDS_FLAGS_BYTES (2u << 20)
extern "C" int gds_system_create(gds_ctx* c, gds_graph* g, int64_t n_state, int32_t scheme, gds_system** out) {
  if (!c || !g || !out || n_state <= 0) GDS_FAIL("bad argument");
  if (scheme != GDS_SCHEME_EULER && scheme != GDS_SCHEME_RK4 && scheme != GDS_SCHEME_MAP) GDS_FAIL("unknown scheme");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  gds_system* s = new gds_system();
  s->ctx = c; s->g = g; s->n_state = n_state; s->scheme = scheme;
  size_t bytes = (size_t)(n_state + 4) * sizeof(double);
  for (int i = 0; i < 2; ++i) {
    GDS_CUDA(cudaMalloc((void**)&s->Y[i], bytes));
    GDS_CUDA(cudaMemsetAsync(s->Y[i], 0, bytes, c->stream));
    s->device_bytes += (int64_t)bytes;
  }
  if (scheme == GDS_SCHEME_RK4) {
    for (int i = 0; i < 2; ++i) {
      GDS_CUDA(cudaMalloc((void**)&s->XS[i], bytes));
      GDS_CUDA(cudaMemsetAsync(s->XS[i], 0, bytes, c->stream));
      s->device_bytes += (int64_t)bytes;
    }
    GDS_CUDA(cudaMalloc((void**)&s->ACC, bytes));
    GDS_CUDA(cudaMemsetAsync(s->ACC, 0, bytes, c->stream));
    s->device_bytes += (int64_t)bytes;
  }
  GDS_CUDA(cudaMalloc((void**)&s->d_counter, 2 * sizeof(int64_t)));   // [step, block ticket of the tail kernel]
  GDS_CUDA(cudaMemsetAsync(s->d_counter, 0, 2 * sizeof(int64_t), c->stream));
  // the flag table is exported through CUDA IPC: it gets a whole 2 MiB allocation of its own, so that its handle can
  // never refer to a block shared with other small allocations
  GDS_CUDA(cudaMalloc((void**)&s->d_flags, GDS_FLAGS_BYTES));
  GDS_CUDA(cudaMemsetAsync(s->d_flags, 0, GDS_FLAGS_BYTES, c->stream));
  GDS_CUDA(cudaMalloc((void**)&s->d_p2p_epoch, 2 * sizeof(int64_t)));
  GDS_CUDA(cudaMemsetAsync(s->d_p2p_epoch, 0, 2 * sizeof(int64_t), c->stream));
  GDS_CUDA(cudaHostAlloc((void**)&s->h_p2p_err, sizeof(int64_t), cudaHostAllocMapped));
  *s->h_p2p_err = 0;
  GDS_CUDA(cudaHostGetDevicePointer((void**)&s->d_p2p_err_host, s->h_p2p_err, 0));
  if (const char* e = getenv("GDS_B200_P2P_TIMEOUT_MS")) {
    const double ms = atof(e);
    if (ms > 0.0) s->p2p_timeout_ns = (uint64_t)(ms * 1e6);
  }
  *out = s;
  return 0;
}

extern "C" int gds_system_destroy(gds_system* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  for (int i = 0; i < 2; ++i) {
    if (s->exec[i]) cudaGraphExecDestroy(s->exec[i]);
    if (s->graph[i]) cudaGraphDestroy(s->graph[i]);
    cudaFree(s->Y[i]);
    cudaFree(s->XS[i]);
    cudaFree(s->Z[i]);
  }
  cudaFree(s->ACC);
  cudaFree(s->d_dir_idx_static); cudaFree(s->d_dir_val_static); cudaFree(s->d_dir_idx_dyn);
  cudaFree(s->d_bc_tab); cudaFree(s->d_steptab); cudaFree(s->d_counter);
  cudaFree(s->d_dir_phase_static); cudaFree(s->d_dir_phase_dyn);
  for (auto& pr : s->peers)
    for (void* m : pr.mapped) if (m) cudaIpcCloseMemHandle(m);
  cudaFree(s->d_flags); cudaFree(s->d_p2p_epoch); cudaFree(s->d_p2p_src); cudaFree(s->d_p2p_dst); cudaFree(s->d_p2p_done);
  if (s->h_p2p_err) cudaFreeHost(s->h_p2p_err);
  cudaFree(s->d_rec); cudaFree(s->d_traj);
  for (int i = 0; i < 2; ++i) {
    cudaFree(s->ens_in[i]); cudaFree(s->ens_out[i]);
    if (s->ens_in_ready[i]) cudaEventDestroy(s->ens_in_ready[i]);
    if (s->ens_in_free[i]) cudaEventDestroy(s->ens_in_free[i]);
    if (s->ens_out_ready[i]) cudaEventDestroy(s->ens_out_ready[i]);
    if (s->ens_out_free[i]) cudaEventDestroy(s->ens_out_free[i]);
  }
  if (s->ens_h2d) cudaStreamDestroy(s->ens_h2d);
  if (s->ens_d2h) cudaStreamDestroy(s->ens_d2h);
  delete s;
  return 0;
}

extern "C" int gds_system_add_pass(gds_system* s, const gds_pass_desc* pass) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("system already finalized");
  PassHost ph;
  GDS_TRY(copy_pass(pass, ph));
  if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(ph);
  if (ph.when != GDS_WHEN_STAGE && ph.when != GDS_WHEN_STEP) GDS_FAIL("unknown `when`");
  int64_t rows = domain_rows(s->g, ph.domain);
  if (rows < 0) GDS_FAIL("unknown domain");
  if (ph.epilogue == GDS_EPI_RHS) {
    if (ph.when != GDS_WHEN_STAGE) GDS_FAIL("an RHS pass runs every stage");
    if (ph.state_offset < 0 || ph.state_offset + rows > s->n_state) GDS_FAIL("RHS rows outside the state vector");
    if (ph.dirichlet && ph.dirichlet->n_rows < rows) GDS_FAIL("Dirichlet mask shorter than the pass domain");
  } else if (ph.epilogue != GDS_EPI_NONE) {
    GDS_FAIL("unknown epilogue");
  }
  for (auto& v : ph.vec) {
    if (ph.when == GDS_WHEN_STEP && v.kind == GDS_VEC_STAGE) GDS_FAIL("a per-step pass cannot read the stage state");
    if ((v.kind == GDS_VEC_STATE || v.kind == GDS_VEC_STAGE) && (v.offset < 0 || v.offset > s->n_state))
      GDS_FAIL("operand offset outside the state vector");
  }
  s->passes.push_back(ph);
  return 0;
}

static int add_special(gds_system* s, int32_t special, int64_t dst, int64_t src, int64_t n) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("system already finalized");
  if (n < 0 || dst < 0 || dst + n > s->n_state || src < 0 || src + n > s->n_state) GDS_FAIL("rows outside the state vector");
  PassHost ph;
  ph.domain = GDS_NODES; ph.when = GDS_WHEN_STAGE;
  ph.epilogue = GDS_EPI_RHS; ph.state_offset = dst; ph.dirichlet = nullptr;
  ph.special = special; ph.sp_dst = dst; ph.sp_src = src; ph.sp_n = n;
  if (special == 1) {
    ph.code = {GDS_OP_PUSH_VEC};
    gds_vecref v; v.kind = GDS_VEC_STAGE; v.buf = nullptr; v.offset = src;
    ph.vec = {v};
  } else {
    ph.code = {GDS_OP_PUSH_IMM};
    ph.imm = {0.0};
  }
  if (!getenv("GDS_B200_NO_FASTPATH")) match_forms(ph);
  s->passes.push_back(ph);
  return 0;
}

extern "C" int gds_system_add_shift(gds_system* s, int64_t dst, int64_t src, int64_t n) { return add_special(s, 1, dst, src, n); }
extern "C" int gds_system_add_hold(gds_system* s, int64_t off, int64_t n) { return add_special(s, 2, off, off, n); }

extern "C" int gds_system_set_dirichlet(gds_system* s, const int64_t* idx, const double* values, int64_t n, int32_t dynamic) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("system already finalized");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  for (int64_t i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= s->n_state) GDS_FAIL("Dirichlet index outside the state vector");
  if (n == 0) return 0;
  if (dynamic) {
    if (s->d_dir_idx_dyn) GDS_FAIL("dynamic Dirichlet set twice");
    GDS_CUDA(cudaMalloc((void**)&s->d_dir_idx_dyn, (size_t)n * 8));
    GDS_CUDA(cudaMemcpy(s->d_dir_idx_dyn, idx, (size_t)n * 8, cudaMemcpyHostToDevice));
    s->n_dir_dyn = n;
  } else {
    if (s->d_dir_idx_static) GDS_FAIL("static Dirichlet set twice");
    if (!values) GDS_FAIL("static Dirichlet needs values");
    GDS_CUDA(cudaMalloc((void**)&s->d_dir_idx_static, (size_t)n * 8));
    GDS_CUDA(cudaMalloc((void**)&s->d_dir_val_static, (size_t)n * 8));
    GDS_CUDA(cudaMemcpy(s->d_dir_idx_static, idx, (size_t)n * 8, cudaMemcpyHostToDevice));
    GDS_CUDA(cudaMemcpy(s->d_dir_val_static, values, (size_t)n * 8, cudaMemcpyHostToDevice));
    s->n_dir_static = n;
  }
  return 0;
}

// per-launch CUDA-event timing for gds_system_step_timed (eager enqueue, not the captured graph)
struct LaunchTimer {
  struct Rec { cudaEvent_t a, b; int cls; };
  std::vector<Rec> recs;
  cudaStream_t stream;
  int begin(int cls) {
    Rec r; r.cls = cls;
    GDS_CUDA(cudaEventCreate(&r.a));
    GDS_CUDA(cudaEventCreate(&r.b));
    GDS_CUDA(cudaEventRecord(r.a, stream));
    recs.push_back(r);
    return 0;
  }
  int end() { GDS_CUDA(cudaEventRecord(recs.back().b, stream)); return 0; }
};
static int domain_class(int32_t domain) { return domain == GDS_NODES ? 0 : (domain == GDS_EDGES ? 1 : 2); }

static int n_stages_of(const gds_system* s) { return (s->scheme == GDS_SCHEME_RK4 && !s->collapsed) ? 4 : 1; }

// per-step passes: sub-expressions of the accepted state only, evaluated once before stage 1
static int enqueue_prestep(gds_system* s, int cur, LaunchTimer* lt, int64_t row0 = 0, int64_t row1 = -1) {
  gds_ctx* c = s->ctx;
  const double* yn = s->Y[cur];
  for (const PassHost& ph : s->passes) {
    if (ph.when != GDS_WHEN_STEP) continue;
    PassParams pp;
    GDS_TRY(resolve_pass(s->g, ph, yn, nullptr, pp));
    pp.row0 = std::min<int64_t>(row0, pp.n_rows);
    if (row1 >= 0) pp.n_rows = std::min<int64_t>(row1, pp.n_rows);
    pp.steptab = s->d_steptab; pp.step_counter = s->d_counter; pp.c_stage = 0.0;
    if (lt) GDS_TRY(lt->begin(domain_class(ph.domain)));
    GDS_TRY(launch_pass(c, pp, ph.domain));
    if (lt) GDS_TRY(lt->end());
  }
  return 0;
}

// vector a stage writes: the next stage state, or the new accepted state after the last stage
static double* stage_output(gds_system* s, int cur, int st) {
  return (s->scheme == GDS_SCHEME_RK4 && !s->collapsed && st < 3) ? s->XS[st & 1] : s->Y[cur ^ 1];
}

// every pass of stage `st` over rows [row0, row1) of each pass's domain (row1 < 0: all rows)
static int enqueue_stage(gds_system* s, int cur, int st, LaunchTimer* lt, int64_t row0 = 0, int64_t row1 = -1) {
  gds_ctx* c = s->ctx;
  const double* yn = s->Y[cur];
  double* ynew = s->Y[cur ^ 1];
  static const double a_next[4] = {0.5, 0.5, 1.0, 0.0};
  static const double c_stage[4] = {0.0, 0.5, 0.5, 1.0};
  static const double acc_w[4] = {1.0, 2.0, 2.0, 1.0};
  const double* stage = (st == 0) ? yn : s->XS[(st - 1) & 1];
  double* xnext = (s->scheme == GDS_SCHEME_RK4 && !s->collapsed && st < 3) ? s->XS[st & 1] : nullptr;
  for (const PassHost& ph : s->passes) {
    if (ph.when != GDS_WHEN_STAGE) continue;
    if (ph.fused_away) continue;   // collapsed stepping: this block shift runs inside the RHS pass of its source block
    PassParams pp;
    GDS_TRY(resolve_pass(s->g, ph, yn, stage, pp));
    if (ph.special) pp.n_rows = ph.sp_n;
    // peer-to-peer halo: ghost rows of the state are written by their owners only
    // (the order-k block shift too: its ghost rows arrive with the derivative block's push like any other)
    if (s->p2p && s->owned_rows >= 0 && ph.epilogue == GDS_EPI_RHS && ph.special != 2)
      pp.n_rows = std::min<int64_t>(s->owned_rows, pp.n_rows);
    pp.row0 = std::min<int64_t>(row0, pp.n_rows);
    if (row1 >= 0) pp.n_rows = std::min<int64_t>(row1, pp.n_rows);
    pp.steptab = s->d_steptab; pp.step_counter = s->d_counter;
    pp.c_stage = (s->scheme == GDS_SCHEME_RK4) ? c_stage[st] : 0.0;
    if (ph.epilogue == GDS_EPI_RHS) {
      int64_t off = ph.state_offset;
      pp.dmask = ph.dirichlet ? ph.dirichlet->d : nullptr;
      pp.yn = yn + off;
      pp.ynew = ynew + off;
      if (s->scheme == GDS_SCHEME_RK4 && s->collapsed) {
        // the law reads accepted state only: all four stage derivatives coincide and the whole step is this one pass
        pp.acc_in = s->ACC + off; pp.acc_out = s->ACC + off;
        pp.fin = 1.0 / 6.0;
        pp.stage_mode = GDS_MODE_COLLAPSED;
        pp.store_k = ph.store_k ? 1 : 0;
        if (ph.fuse_dst >= 0) { pp.yn2 = yn + ph.fuse_dst; pp.ynew2 = ynew + ph.fuse_dst; }
        if (ph.special == 1) {   // order-2 block shift: the row value is the source block's derivative, kept in ACC
          pp.stage_mode = GDS_MODE_COLLAPSED_SHIFT;
          pp.aux = yn + ph.sp_src;
          const double* kv = s->ACC + ph.sp_src;
          pp.vec[0] = kv;
          if (pp.use_terms) pp.lap_b = kv;
        }
      } else if (s->scheme == GDS_SCHEME_RK4) {
        pp.acc_in = s->ACC + off; pp.acc_out = s->ACC + off;
        pp.xnext = xnext ? xnext + off : nullptr;
        pp.acc_w = acc_w[st]; pp.a_next = a_next[st]; pp.fin = 1.0 / 6.0;
        pp.stage_mode = (st == 0) ? 0 : (st == 3 ? 2 : 1);
      } else if (s->scheme == GDS_SCHEME_EULER) {
        pp.stage_mode = 3;
      } else {
        pp.stage_mode = 4;
      }
    }
    if (s->shadow && &ph == &s->passes[(size_t)s->shadow_pass]) {
      // carried gather operand: gather Z = f(Y) (kept beside the state) instead of evaluating f per gathered neighbour;
      // the epilogue writes f of the new state next to the new state
      const int64_t off = ph.state_offset;
      pp.shadow_chain = pp.lap_chain;
      pp.lap_v = s->Z[cur] + off;
      pp.lap_chain.n = 0;
      if (s->shadow_b_same) { pp.lap_b = s->Z[cur] + off; pp.lap_bchain.n = 0; }
      pp.shadow_out = s->Z[cur ^ 1] + off;
    }
    if (lt) GDS_TRY(lt->begin(domain_class(ph.domain)));
    GDS_TRY(launch_pass(c, pp, ph.domain));
    if (lt) GDS_TRY(lt->end());
  }
  return 0;
}

// Dirichlet overwrite (fds.py:289-290,362; coupled: fds.py:486-488), then advance the step counter.
// phase < 0: all of it.  With the overlap split the overwrite of the rows in the boundary ranges (phase 1) runs before
// the push of the new accepted state, the rest (phase 0) and the counter after the interior rows have finished.
static int enqueue_tail(gds_system* s, int cur, LaunchTimer* lt, int phase = -1) {
  gds_ctx* c = s->ctx;
  double* ynew = s->Y[cur ^ 1];
  if (lt) GDS_TRY(lt->begin(3));
  const bool rec = s->recording && !lt;   // the snapshot reads the step counter after the overwrite: advance it last
  const bool last = phase != 1;
  GDS_TRY(launch_tail(c, ynew, s->d_dir_idx_static, s->d_dir_val_static, s->n_dir_static, s->d_dir_idx_dyn, s->d_bc_tab,
                      s->n_dir_dyn, s->d_counter, (rec || !last) ? 0 : 1, s->d_dir_phase_static, s->d_dir_phase_dyn, phase,
                      s->shadow ? s->Z[cur ^ 1] : nullptr, s->shadow ? &s->passes[(size_t)s->shadow_pass].lap.chain : nullptr));
  if (rec && last) {
    GDS_TRY(launch_record(c, ynew, s->d_traj, s->d_rec, s->d_counter, s->n_state));
    GDS_TRY(launch_advance_counter(c, s->d_counter));
  }
  if (lt) GDS_TRY(lt->end());
  return 0;
}

// peer-to-peer halo: push this rank's boundary values of buffer `which` (0/1 = Y, 2/3 = XS) into the neighbours' ghost
// rows (enqueue_push), then wait for theirs (enqueue_wait)
static int enqueue_push(gds_system* s, int which) {
  if (!s->p2p || s->peers.empty()) return 0;
  const double* mine[4] = {s->Y[0], s->Y[1], s->XS[0], s->XS[1]};
  PushParams pu;
  memset(&pu, 0, sizeof(pu));
  pu.n_peers = (int32_t)s->peers.size();
  pu.src = mine[which];
  pu.src_idx = s->d_p2p_src; pu.dst_idx = s->d_p2p_dst;
  pu.epoch = s->d_p2p_epoch; pu.err = s->d_p2p_epoch + 1; pu.err_host = s->d_p2p_err_host;
  pu.timeout_ns = s->p2p_timeout_ns;
  // node-only partitions stop their RHS passes at the owned rows, so nothing but the owners' pushes ever writes a ghost
  // row and no handshake is needed (it costs a second cross-GPU round trip per exchange: 0.32 vs 0.05 ms per RK4 step)
  pu.handshake = s->owned_rows < 0 ? 1 : 0;
  pu.done = s->d_p2p_done;
  int64_t longest = 1;
  for (const gds_system::Peer& pr : s->peers) longest = std::max<int64_t>(longest, pr.send1 - pr.send0);
  pu.blocks_per_peer = (int32_t)std::min<int64_t>(64, std::max<int64_t>(1, (longest + 256 * 8 - 1) / (256 * 8)));
  for (size_t q = 0; q < s->peers.size(); ++q) {
    const gds_system::Peer& pr = s->peers[q];
    pu.dst[q] = (double*)pr.mapped[which];
    pu.flag[q] = (int64_t*)pr.mapped[4] + s->my_rank;
    pu.ready[q] = s->d_flags + 64 + pr.rank;
    pu.s0[q] = pr.send0; pu.s1[q] = pr.send1;
  }
  return launch_push(s->ctx, pu);
}

static int enqueue_wait(gds_system* s) {
  if (!s->p2p || s->peers.empty()) return 0;
  WaitParams wa;
  memset(&wa, 0, sizeof(wa));
  wa.n_peers = (int32_t)s->peers.size();
  wa.epoch = s->d_p2p_epoch; wa.err = s->d_p2p_epoch + 1; wa.err_host = s->d_p2p_err_host;
  wa.timeout_ns = s->p2p_timeout_ns;
  for (size_t q = 0; q < s->peers.size(); ++q) wa.flag[q] = s->d_flags + s->peers[q].rank;
  return launch_wait(s->ctx, wa);
}

static bool overlapped(const gds_system* s) {
  return s->p2p && !s->peers.empty() && s->owned_rows >= 0 && s->ov_lo_end >= 0 && s->ov_hi_begin >= s->ov_lo_end;
}

// One step with the exchange hidden behind the interior rows (peer-to-peer halo + gds_system_p2p_set_overlap).
// Rows [0, lo) and [hi, owned) -- the BOUNDARY ranges -- hold every row that is sent to, or reads a ghost row from, a
// neighbour; rows [lo, hi) -- the INTERIOR -- touch owned rows only.  Two branches, both captured into the step graph:
//
//   main:  ... wait(k-1) -> boundary(st) -> [tail: Dirichlet rows in the boundary ranges] -> push(k) -> wait(k) -> boundary(st+1) ...
//   side:  ... interior(st-1) ----------> interior(st) ------------------------------------------------> interior(st+1) ...
//
// boundary(st) needs interior(st-1) and interior(st) needs boundary(st-1) (rows next to the seam), but the interior never
// waits for an exchange: the critical path of a step is its interior kernels back to back, and the NVLink stores, the
// flag round trip and the skew between ranks run beside them.  A neighbour that is one exchange ahead only stores into
// ghost rows, which no interior row reads.  Per-step passes (sub-expressions of the accepted state) are split the same
// way.  The step counter is advanced on the side branch, after both branches have read this step's table row.
static int enqueue_step_overlapped(gds_system* s, int cur) {
  gds_ctx* c = s->ctx;
  cudaStream_t M = c->stream, S = c->side_stream;
  const int64_t lo = s->ov_lo_end, hi = s->ov_hi_begin;
  auto mark = [&](cudaStream_t st, cudaEvent_t* out) -> int {   // a point on `st` another branch can wait for
    cudaEvent_t e = c->ev_pool[c->ev_next];
    c->ev_next = (c->ev_next + 1) % 16;
    GDS_CUDA(cudaEventRecord(e, st));
    *out = e;
    return 0;
  };
  struct OnStream {   // launches go to ctx->stream: point it at a branch for the lifetime of this object
    gds_ctx* c; cudaStream_t saved;
    OnStream(gds_ctx* c_, cudaStream_t st) : c(c_), saved(c_->stream) { c->stream = st; }
    ~OnStream() { c->stream = saved; }
  };
  const int ns = n_stages_of(s);
  cudaEvent_t evM, evS;
  GDS_TRY(mark(M, &evM));
  GDS_CUDA(cudaStreamWaitEvent(S, evM, 0));               // fork
  {
    OnStream on(c, S);
    GDS_TRY(enqueue_prestep(s, cur, nullptr, lo, hi));
  }
  GDS_TRY(enqueue_prestep(s, cur, nullptr, 0, lo));
  GDS_TRY(enqueue_prestep(s, cur, nullptr, hi, -1));
  GDS_TRY(mark(M, &evM));
  GDS_TRY(mark(S, &evS));
  for (int st = 0; st < ns; ++st) {
    const int which = (st < ns - 1) ? 2 + (st & 1) : (cur ^ 1);   // x_next of this stage / the new accepted state
    const bool last = st == ns - 1;
    GDS_CUDA(cudaStreamWaitEvent(M, evS, 0));             // boundary(st) reads what interior(st-1) wrote
    GDS_CUDA(cudaStreamWaitEvent(S, evM, 0));             // interior(st) reads what boundary(st-1) wrote
    GDS_TRY(enqueue_stage(s, cur, st, nullptr, 0, lo));
    GDS_TRY(enqueue_stage(s, cur, st, nullptr, hi, -1));
    GDS_TRY(mark(M, &evM));                                // before the push: the side branch never waits for an exchange
    {
      OnStream on(c, S);
      GDS_TRY(enqueue_stage(s, cur, st, nullptr, lo, hi));
    }
    GDS_TRY(mark(S, &evS));
    if (last) {
      // Dirichlet overwrite of the new accepted state: rows in the boundary ranges before the push (they may be among
      // the rows sent), the others and the step counter on the side branch once main has read this step's table row
      GDS_TRY(enqueue_tail(s, cur, nullptr, 1));
      cudaEvent_t evT;
      GDS_TRY(mark(M, &evT));
      GDS_CUDA(cudaStreamWaitEvent(S, evT, 0));
      OnStream on(c, S);
      GDS_TRY(enqueue_tail(s, cur, nullptr, 0));
    }
    GDS_TRY(enqueue_push(s, which));
    GDS_TRY(enqueue_wait(s));
  }
  GDS_TRY(mark(S, &evS));
  GDS_CUDA(cudaStreamWaitEvent(M, evS, 0));               // join
  return 0;
}

// enqueue every kernel of one step reading Y[cur], writing Y[cur^1].  Timed stepping (`lt`) keeps one launch per pass
// and runs the exchange between the event pairs.
static int enqueue_step(gds_system* s, int cur, LaunchTimer* lt = nullptr) {
  if (overlapped(s) && !lt) return enqueue_step_overlapped(s, cur);
  GDS_TRY(enqueue_prestep(s, cur, lt));
  const int ns = n_stages_of(s);
  for (int st = 0; st < ns; ++st) {
    const int which = (st < ns - 1) ? 2 + (st & 1) : (cur ^ 1);   // x_next of this stage / the new accepted state
    GDS_TRY(enqueue_stage(s, cur, st, lt));
    if (st == ns - 1) GDS_TRY(enqueue_tail(s, cur, lt));
    GDS_TRY(enqueue_push(s, which));
    GDS_TRY(enqueue_wait(s));
  }
  return 0;
}

#define GDS_STEP_CAPACITY 4096  // steps per table upload; longer runs are chunked

// overlap split: phase of every Dirichlet entry = 1 when its row lies in the boundary ranges of the pass that advances it
static int classify_dirichlet(gds_system* s) {
  auto phase_of = [&](int64_t idx) -> uint8_t {
    for (const PassHost& ph : s->passes) {
      if (ph.epilogue != GDS_EPI_RHS) continue;
      const int64_t rows = ph.special ? ph.sp_n : domain_rows(s->g, ph.domain);
      if (idx < ph.state_offset || idx >= ph.state_offset + rows) continue;
      const int64_t row = idx - ph.state_offset;
      return (row < s->ov_lo_end || row >= s->ov_hi_begin) ? 1 : 0;
    }
    return 0;
  };
  auto build = [&](const int64_t* d_idx, int64_t n, uint8_t** d_out) -> int {
    if (n <= 0) return 0;
    std::vector<int64_t> idx((size_t)n);
    GDS_CUDA(cudaMemcpy(idx.data(), d_idx, (size_t)n * 8, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> ph((size_t)n);
    for (int64_t i = 0; i < n; ++i) ph[i] = phase_of(idx[i]);
    GDS_CUDA(cudaMalloc((void**)d_out, (size_t)n));
    GDS_CUDA(cudaMemcpy(*d_out, ph.data(), (size_t)n, cudaMemcpyHostToDevice));
    return 0;
  };
  GDS_TRY(build(s->d_dir_idx_static, s->n_dir_static, &s->d_dir_phase_static));
  GDS_TRY(build(s->d_dir_idx_dyn, s->n_dir_dyn, &s->d_dir_phase_dyn));
  return 0;
}

extern "C" int gds_system_finalize(gds_system* s) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) return 0;
  gds_ctx* c = s->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  // every state row must be advanced by exactly one RHS pass
  {
    std::vector<std::pair<int64_t, int64_t>> cover;
    for (const PassHost& ph : s->passes)
      if (ph.epilogue == GDS_EPI_RHS)
        cover.emplace_back(ph.state_offset, ph.special ? ph.sp_n : domain_rows(s->g, ph.domain));
    std::sort(cover.begin(), cover.end());
    int64_t at = 0;
    for (auto& iv : cover) {
      if (iv.first != at) GDS_FAIL("RHS passes do not tile the state vector (gap or overlap at " + std::to_string(at) + ")");
      at += iv.second;
    }
    if (at != s->n_state) GDS_FAIL("RHS passes do not cover the state vector");
  }
  if (s->collapsed) {
    // the shifts of an order-2 field read the derivative of their source block from ACC: its RHS pass must run first
    // and store it.  Higher orders chain stage values through two shifts and are not collapsed.
    for (size_t i = 0; i < s->passes.size(); ++i) {
      PassHost& sh = s->passes[i];
      if (sh.special != 1) continue;
      bool found = false;
      for (size_t j = 0; j < s->passes.size(); ++j) {
        PassHost& src = s->passes[j];
        if (src.epilogue != GDS_EPI_RHS || src.special || src.state_offset != sh.sp_src) continue;
        if (j > i) GDS_FAIL("collapsed stepping: a block shift is ordered before the pass it reads");
        // same rows on both sides (the blocks of one field): the shift is evaluated in the epilogue of the source's RHS
        // pass -- no second launch, no derivative written to memory; otherwise the derivative goes through ACC
        if (src.fuse_dst < 0 && sh.sp_n == domain_rows(s->g, src.domain) && !getenv("GDS_B200_NO_FUSED_SHIFT")) {
          src.fuse_dst = sh.sp_dst;
          sh.fused_away = true;
        } else {
          src.store_k = true;
        }
        found = true;
      }
      if (!found) GDS_FAIL("collapsed stepping supports fields of order <= 2 (a block shift reads another shifted block)");
    }
    for (const PassHost& ph : s->passes)
      for (const gds_vecref& v : ph.vec)
        if (v.kind == GDS_VEC_STAGE && !ph.special) GDS_FAIL("collapsed stepping: a pass reads the stage state");
  }
  if (s->shadow) {
    // which pass can gather a carried operand: a map whose one Laplacian pass applies a pointwise chain f to its own
    // stage vector (CML: y <- f(y) + c L f(y)); anything else simply keeps the on-the-fly evaluation
    s->shadow_pass = -1;
    int n_rhs = 0;
    for (size_t i = 0; i < s->passes.size(); ++i) {
      const PassHost& ph = s->passes[i];
      if (ph.epilogue != GDS_EPI_RHS || ph.special == 2) continue;
      ++n_rhs;
      if (ph.special || !ph.lap.ok || ph.lap.chain.n == 0 || ph.domain != GDS_NODES) continue;
      const gds_vecref& v = ph.vec[(size_t)ph.lap.v];
      if (v.kind != GDS_VEC_STAGE || v.offset != ph.state_offset) continue;
      // beta g(b) on the row's own value: read from Z as well when b is the same vector with g = f, otherwise left as it
      // is (one evaluation per row)
      s->shadow_b_same = false;
      if (ph.lap.b >= 0) {
        const gds_vecref& b = ph.vec[(size_t)ph.lap.b];
        bool same = b.kind == v.kind && b.offset == v.offset && ph.lap.bchain.n == ph.lap.chain.n;
        for (int k = 0; same && k < ph.lap.chain.n; ++k)
          same = ph.lap.bchain.pre[k] == ph.lap.chain.pre[k] && ph.lap.bchain.m[k] == ph.lap.chain.m[k] &&
                 ph.lap.bchain.a[k] == ph.lap.chain.a[k];
        s->shadow_b_same = same;
      }
      s->shadow_pass = (int)i;
    }
    if (s->scheme != GDS_SCHEME_MAP || n_rhs != 1 || s->shadow_pass < 0 || s->p2p) {
      s->shadow = false;
    } else {
      const size_t bytes = (size_t)(s->n_state + 4) * sizeof(double);
      for (int i = 0; i < 2; ++i) {
        GDS_CUDA(cudaMalloc((void**)&s->Z[i], bytes));
        GDS_CUDA(cudaMemsetAsync(s->Z[i], 0, bytes, c->stream));
        s->device_bytes += (int64_t)bytes;
      }
    }
  }
  s->cap_steps = GDS_STEP_CAPACITY;
  GDS_CUDA(cudaMalloc((void**)&s->d_steptab, (size_t)s->cap_steps * sizeof(StepRec)));
  GDS_CUDA(cudaMemset(s->d_steptab, 0, (size_t)s->cap_steps * sizeof(StepRec)));
  if (s->n_dir_dyn > 0) GDS_CUDA(cudaMalloc((void**)&s->d_bc_tab, (size_t)s->cap_steps * s->n_dir_dyn * 8));
  if (s->recording) {
    GDS_CUDA(cudaMalloc((void**)&s->d_rec, (size_t)s->cap_steps * 8));
    GDS_CUDA(cudaMemset(s->d_rec, 0xff, (size_t)s->cap_steps * 8));   // -1: record nothing
    GDS_CUDA(cudaMalloc((void**)&s->d_traj, sizeof(double*)));
    GDS_CUDA(cudaMemset(s->d_traj, 0, sizeof(double*)));
  }
  if (c->stream != c->own_stream) GDS_FAIL("finalize on the context's own stream (an external stream cannot be captured here)");
  if (overlapped(s)) GDS_TRY(classify_dirichlet(s));
  for (int cur = 0; cur < 2; ++cur) {
    int64_t before = c->kernel_launches;
    GDS_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_step(s, cur);
    cudaError_t e = cudaStreamEndCapture(c->stream, &s->graph[cur]);
    if (rc) return rc;
    if (e != cudaSuccess) GDS_FAIL(std::string("stream capture failed: ") + cudaGetErrorString(e));
    GDS_CUDA(cudaGraphInstantiate(&s->exec[cur], s->graph[cur], 0));
    s->kernels_per_step = c->kernel_launches - before;
    c->kernel_launches = before;  // captured, not launched
  }
  s->finalized = true;
  return 0;
}

extern "C" int gds_system_set_state(gds_system* s, const double* host, int64_t off, int64_t n) {
  if (!s || off < 0 || n < 0 || off + n > s->n_state) GDS_FAIL("range outside the state vector");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  GDS_CUDA(cudaMemcpyAsync(s->Y[s->cur] + off, host, (size_t)n * 8, cudaMemcpyHostToDevice, s->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return 0;
}

extern "C" int gds_system_get_state(gds_system* s, double* host, int64_t off, int64_t n) {
  if (!s || off < 0 || n < 0 || off + n > s->n_state) GDS_FAIL("range outside the state vector");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  GDS_CUDA(cudaMemcpyAsync(host, s->Y[s->cur] + off, (size_t)n * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return 0;
}

extern "C" int gds_system_state_devptr(gds_system* s, void** dev_out) {
  if (!s || !dev_out) GDS_FAIL("bad argument");
  *dev_out = s->Y[s->cur];
  return 0;
}

static int system_step(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc,
                       const int64_t* rec_slot);

// carried gather operand: Z = f(Y) for the current accepted state.  Done at the start of every stepping call -- between
// two calls the host may have written the state -- and kept up to date by the kernels that write the state afterwards.
static int shadow_refresh(gds_system* s) {
  if (!s->shadow) return 0;
  return launch_shadow_init(s->ctx, s->Y[s->cur], s->Z[s->cur], s->n_state, s->passes[(size_t)s->shadow_pass].lap.chain);
}

extern "C" int gds_system_step(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc) {
  return system_step(s, n_steps, t, h, bc, nullptr);
}

// ---- trajectory recording (System.solve, gds/system.py:38-57: a host copy of every observable after every step there):
//      snapshots are taken on the device inside the captured step and read back once
extern "C" int gds_system_enable_recording(gds_system* s) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("enable recording before gds_system_finalize (the snapshot kernel is part of the captured step)");
  s->recording = true;
  return 0;
}

extern "C" int gds_system_set_trajectory(gds_system* s, gds_buf* traj) {
  if (!s || !traj) GDS_FAIL("bad argument");
  if (!s->recording || !s->finalized) GDS_FAIL("recording is not enabled on this system");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  s->traj_slots = traj->n / s->n_state;
  GDS_CUDA(cudaMemcpyAsync(s->d_traj, &traj->d, sizeof(double*), cudaMemcpyHostToDevice, s->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return 0;
}

extern "C" int gds_system_step_record(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc,
                                      const int64_t* rec_slot) {
  if (!s || !rec_slot) GDS_FAIL("bad argument");
  if (!s->recording || s->traj_slots <= 0) GDS_FAIL("no trajectory buffer set");
  for (int64_t i = 0; i < n_steps; ++i)
    if (rec_slot[i] >= s->traj_slots) GDS_FAIL("snapshot slot outside the trajectory buffer");
  return system_step(s, n_steps, t, h, bc, rec_slot);
}

static int system_step(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc,
                       const int64_t* rec_slot) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (n_steps < 0) GDS_FAIL("negative step count");
  if (s->n_dir_dyn > 0 && !bc) GDS_FAIL("dynamic Dirichlet values missing");
  gds_ctx* c = s->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  std::vector<StepRec> tab;
  int64_t done = 0;
  while (done < n_steps) {
    int64_t m = std::min<int64_t>(n_steps - done, s->cap_steps);
    tab.resize((size_t)m);
    for (int64_t i = 0; i < m; ++i) { tab[i].t = t[done + i]; tab[i].h = h[done + i]; }
    // pageable source: the copy is staged before the call returns, so `tab` may be reused
    GDS_CUDA(cudaMemcpyAsync(s->d_steptab, tab.data(), (size_t)m * sizeof(StepRec), cudaMemcpyHostToDevice, c->stream));
    if (s->n_dir_dyn > 0)
      GDS_CUDA(cudaMemcpyAsync(s->d_bc_tab, bc + done * s->n_dir_dyn, (size_t)m * s->n_dir_dyn * 8,
                               cudaMemcpyHostToDevice, c->stream));
    if (s->recording) {
      if (rec_slot) {
        GDS_CUDA(cudaMemcpyAsync(s->d_rec, rec_slot + done, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
        GDS_CUDA(cudaStreamSynchronize(c->stream));   // pageable source owned by the caller
      } else {
        GDS_CUDA(cudaMemsetAsync(s->d_rec, 0xff, (size_t)m * 8, c->stream));
      }
    }
    GDS_CUDA(cudaMemsetAsync(s->d_counter, 0, sizeof(int64_t), c->stream));
    if (done == 0) GDS_TRY(shadow_refresh(s));
    for (int64_t i = 0; i < m; ++i) {
      GDS_CUDA(cudaGraphLaunch(s->exec[s->cur], c->stream));
      s->cur ^= 1;
      c->graph_replays++;
      c->kernel_launches += s->kernels_per_step;
    }
    done += m;
  }
  return 0;
}

// ---- ensemble stepping with host buffers (include/gds_b200.h): member m + 1 is uploaded and member m - 1 downloaded
//      while member m steps.  Three streams: H2D into IN[b], the context's stream (IN[b] -> Y, the captured steps,
//      Y -> OUT[b]), D2H out of OUT[b]; b = m & 1.  Events order the reuse of the two staging pairs.
static int ensemble_setup(gds_system* s) {
  if (s->ens_h2d) return 0;
  // every resource is created once: a call that failed half-way (out of memory) is resumed, not repeated, by the next one
  for (int i = 0; i < 2; ++i) {
    if (!s->ens_in[i]) { GDS_CUDA(cudaMalloc((void**)&s->ens_in[i], (size_t)s->n_state * 8)); s->device_bytes += s->n_state * 8; }
    if (!s->ens_out[i]) { GDS_CUDA(cudaMalloc((void**)&s->ens_out[i], (size_t)s->n_state * 8)); s->device_bytes += s->n_state * 8; }
    if (!s->ens_in_ready[i]) GDS_CUDA(cudaEventCreateWithFlags(&s->ens_in_ready[i], cudaEventDisableTiming));
    if (!s->ens_in_free[i]) GDS_CUDA(cudaEventCreateWithFlags(&s->ens_in_free[i], cudaEventDisableTiming));
    if (!s->ens_out_ready[i]) GDS_CUDA(cudaEventCreateWithFlags(&s->ens_out_ready[i], cudaEventDisableTiming));
    if (!s->ens_out_free[i]) GDS_CUDA(cudaEventCreateWithFlags(&s->ens_out_free[i], cudaEventDisableTiming));
  }
  if (!s->ens_d2h) GDS_CUDA(cudaStreamCreateWithFlags(&s->ens_d2h, cudaStreamNonBlocking));
  GDS_CUDA(cudaStreamCreateWithFlags(&s->ens_h2d, cudaStreamNonBlocking));      // last: its presence marks the set complete
  return 0;
}

extern "C" int gds_system_step_ensemble(gds_system* s, int64_t n_members, int32_t n_seg, const int64_t* seg_off,
                                        const int64_t* seg_n, const double* const* host_in, double* const* host_out,
                                        int64_t n_steps, const double* t, const double* h, const double* bc) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (s->p2p || s->owned_rows >= 0) GDS_FAIL("ensemble stepping is not available on a rank of a partitioned system");
  if (n_members < 0 || n_seg <= 0 || !seg_off || !seg_n || !host_in || !host_out) GDS_FAIL("bad argument");
  if (n_steps <= 0 || n_steps > s->cap_steps) GDS_FAIL("ensemble stepping takes 1 .. cap_steps steps per member");
  if (s->n_dir_dyn > 0 && !bc) GDS_FAIL("dynamic Dirichlet values missing");
  for (int k = 0; k < n_seg; ++k)
    if (seg_off[k] < 0 || seg_n[k] < 0 || seg_off[k] + seg_n[k] > s->n_state) GDS_FAIL("segment outside the state vector");
  for (int64_t i = 0; i < n_members * n_seg; ++i)
    if (!host_in[i] || !host_out[i]) GDS_FAIL("null host buffer");
  if (n_members == 0) return 0;
  gds_ctx* c = s->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  GDS_TRY(ensemble_setup(s));
  // the step tables are the same for every member: uploaded once
  std::vector<StepRec> tab((size_t)n_steps);
  for (int64_t i = 0; i < n_steps; ++i) { tab[i].t = t[i]; tab[i].h = h[i]; }
  GDS_CUDA(cudaMemcpyAsync(s->d_steptab, tab.data(), (size_t)n_steps * sizeof(StepRec), cudaMemcpyHostToDevice, c->stream));
  if (s->n_dir_dyn > 0)
    GDS_CUDA(cudaMemcpyAsync(s->d_bc_tab, bc, (size_t)n_steps * s->n_dir_dyn * 8, cudaMemcpyHostToDevice, c->stream));
  if (s->recording) GDS_CUDA(cudaMemsetAsync(s->d_rec, 0xff, (size_t)n_steps * 8, c->stream));
  // rows of the state outside the segments keep what the system holds now, for every member
  const size_t state_bytes = (size_t)s->n_state * 8;
  for (int b = 0; b < 2 && b < n_members; ++b) {
    GDS_CUDA(cudaMemcpyAsync(s->ens_in[b], s->Y[s->cur], state_bytes, cudaMemcpyDeviceToDevice, c->stream));
    GDS_CUDA(cudaEventRecord(s->ens_in_free[b], c->stream));
    GDS_CUDA(cudaEventRecord(s->ens_out_free[b], c->stream));
  }
  int rc = 0;
  auto fail = [&](cudaError_t e) { if (e != cudaSuccess && !rc) { gds_set_error(std::string("CUDA: ") + cudaGetErrorString(e)); rc = 1; } };
  for (int64_t m = 0; m < n_members && !rc; ++m) {
    const int b = (int)(m & 1);
    fail(cudaStreamWaitEvent(s->ens_h2d, s->ens_in_free[b], 0));
    for (int k = 0; k < n_seg; ++k)
      fail(cudaMemcpyAsync(s->ens_in[b] + seg_off[k], host_in[m * n_seg + k], (size_t)seg_n[k] * 8, cudaMemcpyHostToDevice, s->ens_h2d));
    fail(cudaEventRecord(s->ens_in_ready[b], s->ens_h2d));
    fail(cudaStreamWaitEvent(c->stream, s->ens_in_ready[b], 0));
    fail(cudaMemcpyAsync(s->Y[s->cur], s->ens_in[b], state_bytes, cudaMemcpyDeviceToDevice, c->stream));
    fail(cudaEventRecord(s->ens_in_free[b], c->stream));
    fail(cudaMemsetAsync(s->d_counter, 0, sizeof(int64_t), c->stream));
    if (!rc) rc = shadow_refresh(s);
    for (int64_t i = 0; i < n_steps && !rc; ++i) {
      fail(cudaGraphLaunch(s->exec[s->cur], c->stream));
      s->cur ^= 1;
      c->graph_replays++;
      c->kernel_launches += s->kernels_per_step;
    }
    fail(cudaStreamWaitEvent(c->stream, s->ens_out_free[b], 0));
    fail(cudaMemcpyAsync(s->ens_out[b], s->Y[s->cur], state_bytes, cudaMemcpyDeviceToDevice, c->stream));
    fail(cudaEventRecord(s->ens_out_ready[b], c->stream));
    fail(cudaStreamWaitEvent(s->ens_d2h, s->ens_out_ready[b], 0));
    for (int k = 0; k < n_seg; ++k)
      fail(cudaMemcpyAsync(host_out[m * n_seg + k], s->ens_out[b] + seg_off[k], (size_t)seg_n[k] * 8, cudaMemcpyDeviceToHost, s->ens_d2h));
    fail(cudaEventRecord(s->ens_out_free[b], s->ens_d2h));
  }
  cudaError_t e0 = cudaStreamSynchronize(s->ens_h2d), e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(s->ens_d2h);
  if (rc) return rc;
  if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess)
    GDS_FAIL(std::string("ensemble stepping failed: ") + cudaGetErrorString(e0 != cudaSuccess ? e0 : e1 != cudaSuccess ? e1 : e2));
  return 0;
}

// Same work as gds_system_step, enqueued kernel by kernel with a CUDA-event pair around every launch (no graph
// replay), so that the per-kernel device time is measured live.  class 0/1/2: passes over nodes/edges/faces,
// class 3: the per-step tail (Dirichlet overwrite + step counter).  Synchronises.
extern "C" int gds_system_step_timed(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc,
                                     double* class_ms, int64_t* class_launches) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (n_steps < 0 || n_steps > 256) GDS_FAIL("timed stepping takes at most 256 steps per call");
  if (n_steps > s->cap_steps) GDS_FAIL("too many steps");
  if (s->n_dir_dyn > 0 && !bc) GDS_FAIL("dynamic Dirichlet values missing");
  gds_ctx* c = s->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  std::vector<StepRec> tab((size_t)n_steps);
  for (int64_t i = 0; i < n_steps; ++i) { tab[i].t = t[i]; tab[i].h = h[i]; }
  GDS_CUDA(cudaMemcpyAsync(s->d_steptab, tab.data(), (size_t)n_steps * sizeof(StepRec), cudaMemcpyHostToDevice, c->stream));
  if (s->n_dir_dyn > 0)
    GDS_CUDA(cudaMemcpyAsync(s->d_bc_tab, bc, (size_t)n_steps * s->n_dir_dyn * 8, cudaMemcpyHostToDevice, c->stream));
  GDS_CUDA(cudaMemsetAsync(s->d_counter, 0, sizeof(int64_t), c->stream));
  GDS_TRY(shadow_refresh(s));
  LaunchTimer lt;
  lt.stream = c->stream;
  int rc = 0;
  for (int64_t i = 0; i < n_steps && !rc; ++i) {
    rc = enqueue_step(s, s->cur, &lt);
    s->cur ^= 1;
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  for (int k = 0; k < 4; ++k) { if (class_ms) class_ms[k] = 0.0; if (class_launches) class_launches[k] = 0; }
  for (auto& r : lt.recs) {
    float ms = 0.f;
    if (!rc && e == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (class_ms) class_ms[r.cls] += (double)ms;
      if (class_launches) class_launches[r.cls] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (rc) return rc;
  if (e != cudaSuccess) GDS_FAIL(std::string("stream sync failed: ") + cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// stage-granular stepping (one rank of a partitioned system: the halo exchange sits between the stages)
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_system_begin_steps(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (n_steps < 0 || n_steps > s->cap_steps) GDS_FAIL("step count outside the table capacity (4096)");
  if (s->n_dir_dyn > 0 && !bc) GDS_FAIL("dynamic Dirichlet values missing");
  gds_ctx* c = s->ctx;
  GDS_CUDA(cudaSetDevice(c->device));
  std::vector<StepRec> tab((size_t)n_steps);
  for (int64_t i = 0; i < n_steps; ++i) { tab[i].t = t[i]; tab[i].h = h[i]; }
  GDS_CUDA(cudaMemcpyAsync(s->d_steptab, tab.data(), (size_t)n_steps * sizeof(StepRec), cudaMemcpyHostToDevice, c->stream));
  if (s->n_dir_dyn > 0)
    GDS_CUDA(cudaMemcpyAsync(s->d_bc_tab, bc, (size_t)n_steps * s->n_dir_dyn * 8, cudaMemcpyHostToDevice, c->stream));
  GDS_CUDA(cudaMemsetAsync(s->d_counter, 0, sizeof(int64_t), c->stream));
  GDS_TRY(shadow_refresh(s));
  GDS_CUDA(cudaStreamSynchronize(c->stream));  // the tables come from pageable memory the caller may reuse
  return 0;
}

extern "C" int gds_system_run_stage(gds_system* s, int32_t stage, int64_t row_begin, int64_t row_end) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (stage < 0 || stage >= n_stages_of(s)) GDS_FAIL("no such stage");
  if (row_begin < 0 || (row_begin & 31)) GDS_FAIL("row_begin must be a non-negative multiple of 32");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  if (stage == 0 && row_begin == 0) GDS_TRY(enqueue_prestep(s, s->cur, nullptr));
  return enqueue_stage(s, s->cur, stage, nullptr, row_begin, row_end);
}

extern "C" int gds_system_stage_output(gds_system* s, int32_t stage, void** dev_out) {
  if (!s || !dev_out) GDS_FAIL("bad argument");
  if (stage < 0 || stage >= n_stages_of(s)) GDS_FAIL("no such stage");
  *dev_out = stage_output(s, s->cur, stage);
  return 0;
}

extern "C" int gds_system_end_step(gds_system* s) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  GDS_TRY(enqueue_tail(s, s->cur, nullptr));
  s->cur ^= 1;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// peer-to-peer halo over NVLink
// ---------------------------------------------------------------------------------------------------
extern "C" int gds_system_ipc_export(gds_system* s, uint8_t* handles) {
  if (!s || !handles) GDS_FAIL("bad argument");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  void* bufs[5] = {s->Y[0], s->Y[1], s->XS[0], s->XS[1], s->d_flags};
  memset(handles, 0, 5 * 64);
  for (int i = 0; i < 5; ++i) {
    if (!bufs[i]) continue;   // Euler / map systems have no stage buffers
    cudaIpcMemHandle_t h;
    GDS_CUDA(cudaIpcGetMemHandle(&h, bufs[i]));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handles + 64 * i, &h, 64);
  }
  return 0;
}

extern "C" int gds_system_p2p_attach(gds_system* s, int32_t my_rank, int32_t n_peers, const int32_t* peer_rank,
                                     const uint8_t* peer_handles, const int64_t* send_ptr, const int64_t* src_idx,
                                     const int64_t* dst_idx, int64_t owned_rows) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("attach the peers before gds_system_finalize (the exchange is part of the captured step)");
  if (n_peers < 0 || n_peers > GDS_MAX_PEERS) GDS_FAIL("at most 8 neighbouring ranks");
  if (my_rank < 0 || my_rank >= 64) GDS_FAIL("rank outside the flag table");
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  s->peers.clear();
  void* mine[5] = {s->Y[0], s->Y[1], s->XS[0], s->XS[1], s->d_flags};
  for (int q = 0; q < n_peers; ++q) {
    gds_system::Peer pr;
    pr.rank = peer_rank[q];
    if (pr.rank < 0 || pr.rank >= 64 || pr.rank == my_rank) GDS_FAIL("bad peer rank");
    pr.send0 = send_ptr[q]; pr.send1 = send_ptr[q + 1];
    for (int i = 0; i < 5; ++i) {
      if (!mine[i]) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, peer_handles + (size_t)(q * 5 + i) * 64, 64);
      GDS_CUDA(cudaIpcOpenMemHandle(&pr.mapped[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    s->peers.push_back(pr);
  }
  const int64_t n = n_peers ? send_ptr[n_peers] : 0;
  GDS_CUDA(cudaMalloc((void**)&s->d_p2p_src, (size_t)std::max<int64_t>(n, 1) * 8));
  GDS_CUDA(cudaMalloc((void**)&s->d_p2p_dst, (size_t)std::max<int64_t>(n, 1) * 8));
  GDS_CUDA(cudaMalloc((void**)&s->d_p2p_done, GDS_MAX_PEERS * sizeof(unsigned int)));
  GDS_CUDA(cudaMemset(s->d_p2p_done, 0, GDS_MAX_PEERS * sizeof(unsigned int)));
  if (n) {
    GDS_CUDA(cudaMemcpy(s->d_p2p_src, src_idx, (size_t)n * 8, cudaMemcpyHostToDevice));
    GDS_CUDA(cudaMemcpy(s->d_p2p_dst, dst_idx, (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  s->my_rank = my_rank;
  s->owned_rows = owned_rows;
  s->p2p = true;
  return 0;
}

// 0 = fine; k > 0: the wait for neighbour number k-1 timed out (a peer stopped): the state is not to be trusted
extern "C" int gds_system_p2p_status(gds_system* s, int64_t* exchanges, int64_t* error) {
  if (!s) GDS_FAIL("null system");
  int64_t h[2] = {0, 0};
  GDS_CUDA(cudaSetDevice(s->ctx->device));
  GDS_CUDA(cudaMemcpyAsync(h, s->d_p2p_epoch, 16, cudaMemcpyDeviceToHost, s->ctx->stream));
  GDS_CUDA(cudaStreamSynchronize(s->ctx->stream));
  if (exchanges) *exchanges = h[0];
  if (error) *error = h[1];
  return 0;
}

// recurrence maps of the form y <- alpha L f(y) + beta f(y) + ... : keep Z = f(Y) beside the state
extern "C" int gds_system_set_shadow(gds_system* s, int32_t on) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("choose the carried gather operand before gds_system_finalize");
  s->shadow = on != 0;
  return 0;
}

// RK4 systems whose laws read accepted state only (no stage-state operand, no stage time): one pass per step
extern "C" int gds_system_set_collapsed(gds_system* s, int32_t on) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("choose collapsed stepping before gds_system_finalize");
  if (on && s->scheme != GDS_SCHEME_RK4) GDS_FAIL("collapsed stepping is a property of the RK4 scheme");
  s->collapsed = on != 0;
  return 0;
}

// the same error flag read from its mirror in mapped host memory: no stream synchronisation, so the host can look before
// every call; it shows failures of work that has completed
extern "C" int gds_system_p2p_error(gds_system* s, int64_t* error) {
  if (!s || !error) GDS_FAIL("bad argument");
  *error = s->h_p2p_err ? *reinterpret_cast<volatile int64_t*>(s->h_p2p_err) : 0;
  return 0;
}

extern "C" int gds_system_p2p_set_overlap(gds_system* s, int64_t lo_end, int64_t hi_begin) {
  if (!s) GDS_FAIL("null system");
  if (s->finalized) GDS_FAIL("set the overlap split before gds_system_finalize");
  if (!s->p2p || s->owned_rows < 0) GDS_FAIL("the overlap split needs a peer-to-peer halo whose owned rows come first");
  if (lo_end < 0 || hi_begin < lo_end || hi_begin > s->owned_rows || (lo_end & 127) || (hi_begin & 127))
    GDS_FAIL("need 0 <= lo_end <= hi_begin <= owned_rows, both multiples of 128");
  s->ov_lo_end = lo_end; s->ov_hi_begin = hi_begin;
  return 0;
}

// halo lists: a set of state-vector positions packed into / scattered from a contiguous exchange buffer
struct gds_halo {
  gds_ctx* ctx;
  int64_t* d_idx = nullptr;
  int64_t n = 0;
};

extern "C" int gds_halo_create(gds_ctx* c, const int64_t* state_idx, int64_t n, gds_halo** out) {
  if (!c || !out || n < 0 || (n > 0 && !state_idx)) GDS_FAIL("bad argument");
  GDS_NEED_DEVICE(c);
  GDS_CUDA(cudaSetDevice(c->device));
  gds_halo* hl = new gds_halo();
  hl->ctx = c; hl->n = n;
  GDS_CUDA(cudaMalloc((void**)&hl->d_idx, (size_t)std::max<int64_t>(n, 1) * 8));
  if (n) GDS_CUDA(cudaMemcpy(hl->d_idx, state_idx, (size_t)n * 8, cudaMemcpyHostToDevice));
  *out = hl;
  return 0;
}

extern "C" int gds_halo_destroy(gds_halo* hl) {
  if (!hl) return 0;
  cudaSetDevice(hl->ctx->device);
  cudaStreamSynchronize(hl->ctx->stream);
  cudaFree(hl->d_idx);
  delete hl;
  return 0;
}

extern "C" int gds_halo_pack(gds_halo* hl, const void* vec_dev, void* buf_dev) {
  if (!hl || !vec_dev || !buf_dev) GDS_FAIL("bad argument");
  GDS_CUDA(cudaSetDevice(hl->ctx->device));
  return launch_halo(hl->ctx, (double*)vec_dev, (double*)buf_dev, hl->d_idx, hl->n, 0);
}

extern "C" int gds_halo_unpack(gds_halo* hl, const void* buf_dev, void* vec_dev) {
  if (!hl || !vec_dev || !buf_dev) GDS_FAIL("bad argument");
  GDS_CUDA(cudaSetDevice(hl->ctx->device));
  return launch_halo(hl->ctx, (double*)vec_dev, (double*)buf_dev, hl->d_idx, hl->n, 1);
}

extern "C" int gds_system_modes(gds_system* s, int32_t* collapsed, int32_t* shadow, int32_t* overlap) {
  if (!s) GDS_FAIL("null system");
  if (!s->finalized) GDS_FAIL("system not finalized");
  if (collapsed) *collapsed = s->collapsed ? 1 : 0;
  if (shadow) *shadow = s->shadow ? 1 : 0;
  if (overlap) *overlap = overlapped(s) ? 1 : 0;
  return 0;
}

extern "C" int gds_system_info(gds_system* s, int64_t* kps, int64_t* bytes) {
  if (!s) GDS_FAIL("null system");
  if (kps) *kps = s->kernels_per_step;
  if (bytes) *bytes = s->device_bytes;
  return 0;
}
