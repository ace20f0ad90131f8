// Internal structures of the gds_b200 stepping core (host + device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "gds_b200.h"

#define GDS_SLICE 32  // rows per sliced-ELL slice = one warp

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
void gds_set_error(const std::string& msg);
#define GDS_FAIL(msg)                                                                        \
  do {                                                                                       \
    gds_set_error(std::string(__func__) + ": " + (msg));                                     \
    return 1;                                                                                \
  } while (0)
#define GDS_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      gds_set_error(std::string(__func__) + ": " #call " -> " + cudaGetErrorString(e__));    \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)
#define GDS_TRY(call)                                                                        \
  do {                                                                                       \
    int r__ = (call);                                                                        \
    if (r__) return r__;                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// sliced-ELL (slice = 32 rows = one warp; column-major inside a slice so that entry j of the 32 rows of a
// slice is one contiguous 128-byte (int32) / 256-byte (double) line).  soff[s] = first entry of slice s in
// units of 32 entries; width of slice s = soff[s+1]-soff[s].  Padding entries hold -1.
// ---------------------------------------------------------------------------------------------------
struct HostSell {
  int64_t n_rows = 0;
  std::vector<uint32_t> soff;  // n_slices + 1
  int64_t n_slots = 0;         // = soff.back() * 32
};

// Host-side lowered graph (index maps kept for export / parity checks)
struct HostGraph {
  int64_t V = 0, E = 0, F = 0;
  std::vector<int32_t> eu, ev;
  std::vector<double> w, omega;
  bool unit_weights = true;
  // B1 rows: node -> incident (edge, sign), sorted by edge id
  std::vector<int64_t> n_ptr;
  std::vector<int32_t> n_edge, n_nbr;
  std::vector<int8_t> n_sign;
  std::vector<double> n_dinv;  // 1/(weighted degree - 1), inf -> 0 (gds.py:253-257)
  // B2 rows: face -> (edge, sign) sorted by edge id ; B2 cols: edge -> (face, sign)
  std::vector<int64_t> f_ptr;
  std::vector<int32_t> f_edge;
  std::vector<int8_t> f_sign;
  std::vector<int64_t> ef_ptr;
  std::vector<int32_t> ef_face;
  std::vector<int8_t> ef_sign;
};

int lower_graph_host(int64_t V, int64_t E, const int32_t* eu, const int32_t* ev, const double* w, int64_t F,
                     const int64_t* face_ptr, const int32_t* face_nodes, HostGraph& out);

// Device view handed to kernels (by value inside PassParams)
struct GraphDev {
  int64_t V, E, F;
  // node incidence, sliced-ELL over nodes
  const uint32_t* n_soff;
  const int32_t* n_nbr;   // neighbour node, padding = own row
  const int16_t* n_nbr16; // same slots as (neighbour - own row) in 16 bits (lattices up to 32767 wide, locally numbered
                          // graphs): half the index bytes for the quad Laplacian kernel.  -32768 = escape: the difference
                          // does not fit, read n_nbr.  Null when more than 1 slot in 64 would escape
  int32_t n_uniw;         // > 0: every slice of the node ELL has this width (slice s starts at slot s * n_uniw * 32)
  int32_t n_quad_odd;     // 1: most quads of the node ELL gather RUNS at offsets that are neither multiples of 4 nor +-1
                          // (a lattice whose row / plane length is odd: 693^3).  The quad kernel then falls back to four
                          // scalar gathers per slot, each touching 8 lines per warp; lapT (lane = row) reads the same run
                          // as one contiguous 256-byte access and is the faster kernel (measured 0.39 vs 0.51 ms)
  const int32_t* n_edge;  // (edge << 1) | (1 if this node is the head v of the edge), padding = -1
  const double* n_w;      // w_e, padding = 0
  const double* n_dinv;   // [V]
  // edges
  const int32_t* e_u;
  const int32_t* e_v;
  const double* e_omega;  // sqrt(w)
  const double* e_w;
  // edge -> faces, sliced-ELL over edges: (face << 1) | (1 if B2[f,e] < 0), padding = -1
  const uint32_t* ef_soff;
  const int32_t* ef_face;
  int32_t ef_maxw;        // widest slice of the edge -> face table (2 on planar graphs)
  // face -> edges, sliced-ELL over faces: (edge << 1) | (1 if B2[f,e] < 0), padding = -1
  const uint32_t* fe_soff;
  const int32_t* fe_edge;
};

// A row program recognised as  value = alpha * Z_mask(L0 v)[row] + beta * b[row] + gamma  (the node-Laplacian laws:
// heat, wave, coupled-map lattices, diffusion terms) runs in the specialised 4-rows-per-thread kernel instead of the
// row-program interpreter.
// a chain of <= GDS_MAX_CHAIN pointwise steps applied to a value on the fly.  Every GDS_CHAIN_* step is stored in the
// branch-free form  x <- pre(x) * m + a  with pre = identity / square / abs: multiplying by 1 and adding 0 are exact, so
// the result is bit-identical to the step written out (x*x, x*k, x+k, k-x = (-1)x + k, -x, |x|, (x*x)*x).
struct Chain {  // plain data: lives inside kernel parameters
  int32_t n;
  int32_t pre[GDS_MAX_CHAIN];  // 0 identity, 1 square, 2 abs, 3 cube
  double m[GDS_MAX_CHAIN];
  double a[GDS_MAX_CHAIN];
};
inline bool chain_push(Chain& c, int code, double k) {
  if (c.n >= GDS_MAX_CHAIN) return false;
  int i = c.n;
  c.pre[i] = 0; c.m[i] = 1.0; c.a[i] = 0.0;
  switch (code) {
    case GDS_CHAIN_SQUARE: c.pre[i] = 1; break;
    case GDS_CHAIN_MUL: c.m[i] = k; break;
    case GDS_CHAIN_ADD: c.a[i] = k; break;
    case GDS_CHAIN_RSUB: c.m[i] = -1.0; c.a[i] = k; break;
    case GDS_CHAIN_NEG: c.m[i] = -1.0; break;
    case GDS_CHAIN_ABS: c.pre[i] = 2; break;
    case GDS_CHAIN_CUBE: c.pre[i] = 3; break;
    default: return false;
  }
  c.n++;
  return true;
}

struct LapForm {
  bool ok = false;
  int v = -1, mask = -1, b = -1, b2 = -1, store = -1;  // operand slots of the pass (-1: absent)
  double alpha = 1.0, beta = 0.0, beta2 = 0.0, gamma = 0.0;  // value = alpha Z(L0 f(v)) + beta g(b) + beta2 b2 + gamma
  Chain chain = {};   // applied to the gathered vector: L0 f(v)
  Chain bchain = {};  // applied to the additive vector: beta * g(b)
};

// More generally, a row program that is a LINEAR COMBINATION of masked one-hop operators,
//   value = sum_k coef_k * Z_{m0_k} Z_{m1_k}( op_k(a_k [, b_k]) )[row] + beta * b[row] + gamma,
// (Hodge Laplacians, advection + diffusion, Navier-Stokes edge law, div / curl intermediates) runs in `terms_kernel`,
// which evaluates the terms one after another without the stack machine.
#define GDS_MAX_TERMS 6
struct TermSym {
  int op = 0, a = -1, b = -1, m0 = -1, m1 = -1;  // opcode, operand slots, mask slots
  double coef = 0.0;
  Chain chain = {};  // N_LAP only
};
struct TermForm {
  bool ok = false;
  int n = 0;
  TermSym t[GDS_MAX_TERMS];
  int b = -1, b2 = -1, store = -1;
  double beta = 0.0, beta2 = 0.0, gamma = 0.0;
  Chain bchain = {};
};

#define GDS_MODE_COLLAPSED 5
#define GDS_MODE_COLLAPSED_SHIFT 6

struct StepRec {  // one row of the per-step table
  double t;       // step start time
  double h;       // step size
};

// Parameters of one pass launch (passed by value: lives in the constant bank, warp-uniform reads)
struct PassParams {
  GraphDev g;
  int64_t row0;    // first row of this launch (multiple of 32)
  int64_t n_rows;  // one past the last row
  int32_t n_code;
  uint32_t code[GDS_MAX_CODE];
  double imm[GDS_MAX_IMM];
  const double* vec[GDS_MAX_VEC];
  const uint32_t* mask[GDS_MAX_MASK];
  double* out[GDS_MAX_OUT];
  // epilogue
  int32_t epilogue;   // GDS_EPI_*
  int32_t stage_mode; // 0 first, 1 middle, 2 last (with acc), 3 last (no acc: Euler), 4 map,
                      // 5 whole RK4 step of a stage-independent law, 6 the same for an order-2 block shift (see GDS_MODE_*)
  const double* aux;  // mode 6: accepted state of the SOURCE block (v_n), while the row value is its derivative k_v
  int32_t store_k;    // mode 5: also write the (Dirichlet-masked) derivative to acc_out, for a block shift that needs it
  // carried gather operand (gds_system_set_shadow): the pass gathers Z = f(state) kept beside the state instead of
  // evaluating f for every gathered neighbour; whoever writes the new state writes f of it as well
  double* shadow_out;
  Chain shadow_chain;
  const double* yn2;  // mode 5, order-2 fields: accepted state of the block advanced by dy/dt = stage value of THIS block
  double* ynew2;      //         ... and where its new value goes: the block shift is fused into this pass (no extra launch)
  const uint32_t* dmask;
  const double* yn;
  const double* acc_in;
  double* acc_out;
  double* xnext;
  double* ynew;
  double acc_w;    // weight of k in acc
  double a_next;   // xnext = yn + a_next*h*k
  double fin;      // ynew = yn + fin*h*(acc + k)
  double c_stage;  // stage time = t + c_stage*h
  const StepRec* steptab;
  const int64_t* step_counter;
  double t_fixed;  // used when steptab == nullptr (eager application)
  // specialised Laplacian form (LapForm resolved to pointers); lap_v == nullptr: run the interpreter
  const double* lap_v;
  const uint32_t* lap_mask;
  const double* lap_b;
  const double* lap_b2;   // second additive vector (e.g. a Neumann source next to a hoisted term)
  double lap_beta2;
  double* lap_store;
  double lap_alpha, lap_beta, lap_gamma;
  Chain lap_chain, lap_bchain;
  // linear combination of masked operators (TermForm resolved to pointers); n_terms == 0: not used.  The additive
  // vector / constant / STORE target reuse lap_b, lap_beta, lap_gamma, lap_store.
  int32_t use_terms;  // 1: run terms_kernel (n_terms may be 0: value = beta * b + gamma, a pointwise pass)
  int32_t n_terms;
  struct {
    int32_t op;
    const double* a;
    const double* b;
    const uint32_t* m0;
    const uint32_t* m1;
    double coef;
    Chain chain;
  } term[GDS_MAX_TERMS];
  int32_t unit_weights;
};

// ---------------------------------------------------------------------------------------------------
// objects behind the opaque handles
// ---------------------------------------------------------------------------------------------------
struct gds_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // second stream + fork / join events: the interior rows of a stage run beside the boundary rows and their push
  // (gds_system_p2p_set_overlap); both are captured into the step graph as two branches
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_pool[16] = {};   // fork / join points of the two branches (used round-robin while a step is enqueued)
  int ev_next = 0;
  int64_t kernel_launches = 0;
  int64_t graph_replays = 0;
  void* flush_buf = nullptr;
  int64_t flush_bytes = 0;
  double* reduce_scratch = nullptr;   // partial results of gds_buf_reduce (fixed grid, fixed tree)
};

struct gds_buf {
  gds_ctx* ctx;
  double* d = nullptr;
  int64_t n = 0;
  bool owned = true;  // false: a view of memory owned elsewhere (gds_buf_view)
};

struct gds_mask {
  gds_ctx* ctx;
  uint32_t* d = nullptr;
  int64_t n_rows = 0;
  int64_t n_set = 0;
};

struct gds_graph {
  gds_ctx* ctx;
  HostGraph h;
  GraphDev d;
  std::vector<void*> allocs;
  int64_t device_bytes = 0;
};

struct PassHost {  // a pass as given by the host, operands still symbolic
  int32_t domain, when;
  std::vector<uint32_t> code;
  std::vector<double> imm;
  std::vector<gds_vecref> vec;
  std::vector<gds_mask*> mask;
  std::vector<gds_buf*> out;
  std::vector<int64_t> out_offset;
  int32_t epilogue;
  int64_t state_offset;
  gds_mask* dirichlet;
  LapForm lap;
  TermForm terms;
  // shift / hold pseudo-passes
  int32_t special = 0;  // 0 normal, 1 shift, 2 hold
  int64_t sp_dst = 0, sp_src = 0, sp_n = 0;
  bool store_k = false; // collapsed stepping: a block shift reads this pass's derivative
  int64_t fuse_dst = -1;  // collapsed stepping: state offset of the shifted block this RHS pass advances as well
  bool fused_away = false;  // collapsed stepping: this block shift runs inside the RHS pass of its source block
};

struct gds_system {
  gds_ctx* ctx;
  gds_graph* g;
  int64_t n_state = 0;
  int32_t scheme = GDS_SCHEME_RK4;
  bool shadow = false;                // carried gather operand Z = f(Y) (gds_system_set_shadow)
  int shadow_pass = -1;               // index of the pass that gathers it
  bool shadow_b_same = false;         // its additive term beta g(b) is beta f(same vector): read from Z too
  double* Z[2] = {nullptr, nullptr};  // f(Y[i])
  bool collapsed = false;             // RK4 of a law that reads accepted state only: one pass per step (gds_system_set_collapsed)
  double* Y[2] = {nullptr, nullptr};  // accepted state ping-pong
  double* XS[2] = {nullptr, nullptr}; // stage state ping-pong
  double* ACC = nullptr;
  int cur = 0;                        // Y[cur] is the accepted state
  std::vector<PassHost> passes;
  // Dirichlet overwrite
  int64_t n_dir_static = 0, n_dir_dyn = 0;
  int64_t* d_dir_idx_static = nullptr;
  double* d_dir_val_static = nullptr;
  int64_t* d_dir_idx_dyn = nullptr;
  uint8_t* d_dir_phase_static = nullptr;  // overlap split: 1 = the entry's row is in the boundary ranges
  uint8_t* d_dir_phase_dyn = nullptr;
  double* d_bc_tab = nullptr;   // [cap_steps x n_dir_dyn]
  StepRec* d_steptab = nullptr; // [cap_steps]
  int64_t cap_steps = 0;
  int64_t* d_counter = nullptr;
  bool finalized = false;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};  // one per parity of `cur`
  cudaGraph_t graph[2] = {nullptr, nullptr};
  int64_t kernels_per_step = 0;
  int64_t device_bytes = 0;
  // trajectory recording (System.solve): after a step whose row of d_rec holds a slot >= 0 the new state is copied into
  // that slot of the trajectory buffer, inside the captured step
  bool recording = false;
  double** d_traj = nullptr;      // device cell holding the trajectory buffer's address (set any time)
  int64_t traj_slots = 0;
  int64_t* d_rec = nullptr;       // [cap_steps]
  // peer-to-peer halo (one rank of a partitioned system): see gds_system_p2p_attach
  struct Peer {
    int rank = -1;
    void* mapped[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // peer's Y0 Y1 XS0 XS1 FLAGS (IPC-mapped)
    int64_t send0 = 0, send1 = 0;                                      // range in d_p2p_src / d_p2p_dst
  };
  std::vector<Peer> peers;
  bool p2p = false;
  int my_rank = 0;
  int64_t owned_rows = -1;        // RHS passes stop here: ghost rows are written by the peers only
  int64_t* d_flags = nullptr;     // [128] arrival flags (slot = sender's rank) and ready flags (64 + rank), written by peers
  int64_t* d_p2p_epoch = nullptr; // [2]: exchanges sent so far; error flag of the wait kernel
  int64_t* d_p2p_src = nullptr;
  int64_t* d_p2p_dst = nullptr;
  unsigned int* d_p2p_done = nullptr;   // per-neighbour block counters of the push kernel
  // the error flag again in pinned, device-mapped host memory: the host polls it without synchronising the stream
  int64_t* h_p2p_err = nullptr;   // host address
  int64_t* d_p2p_err_host = nullptr;  // the same word as seen from the device
  uint64_t p2p_timeout_ns = 20000000000ull;
  // overlap: rows [0, ov_lo_end) and [ov_hi_begin, owned_rows) hold every row that is sent to / reads from a neighbour;
  // they run first, their push follows at once, the interior rows run on a second stream meanwhile (-1: no overlap)
  int64_t ov_lo_end = -1, ov_hi_begin = -1;
  // ensemble stepping with host buffers (gds_system_step_ensemble): double-buffered staging, one stream per PCIe direction
  double* ens_in[2] = {nullptr, nullptr};
  double* ens_out[2] = {nullptr, nullptr};
  cudaStream_t ens_h2d = nullptr, ens_d2h = nullptr;
  cudaEvent_t ens_in_ready[2] = {}, ens_in_free[2] = {}, ens_out_ready[2] = {}, ens_out_free[2] = {};
};

#define GDS_MAX_PEERS 8
struct PushParams {
  int32_t n_peers;
  const double* src;                 // my stage output
  double* dst[GDS_MAX_PEERS];        // the same logical buffer on each peer
  int64_t* flag[GDS_MAX_PEERS];      // peer's arrival flag for this rank (its ready flag is 64 slots further)
  const int64_t* ready[GDS_MAX_PEERS];  // this rank's ready flag written by each peer
  int64_t* err;
  int64_t* err_host;                 // mirror of the error flag in mapped host memory (polled by the host)
  uint64_t timeout_ns;
  int32_t handshake;                 // wait for the receiver's ready flag before storing (deep partitions only)
  int64_t s0[GDS_MAX_PEERS], s1[GDS_MAX_PEERS];
  const int64_t* src_idx;
  const int64_t* dst_idx;
  int64_t* epoch;
  unsigned int* done;                // [GDS_MAX_PEERS] blocks that have finished their stores (self-resetting)
  int32_t blocks_per_peer;
};
struct WaitParams {
  int32_t n_peers;
  const int64_t* flag[GDS_MAX_PEERS];  // my flags, one per neighbour
  const int64_t* epoch;
  int64_t* err;
  int64_t* err_host;
  uint64_t timeout_ns;
};
int launch_record(gds_ctx* ctx, const double* y, double* const* traj, const int64_t* rec, const int64_t* counter, int64_t n);
int launch_push(gds_ctx* ctx, const PushParams& p);
int launch_wait(gds_ctx* ctx, const WaitParams& p);

// kernels.cu
int launch_pass(gds_ctx* ctx, const PassParams& p, int32_t domain);
int launch_dirichlet(gds_ctx* ctx, double* y, const int64_t* idx, const double* val_static, const double* bc_tab,
                     int64_t n, int64_t n_dyn_stride, const int64_t* counter);
int launch_advance_counter(gds_ctx* ctx, int64_t* counter);
int launch_tail(gds_ctx* ctx, double* y, const int64_t* sidx, const double* sval, int64_t ns, const int64_t* didx,
                const double* tab, int64_t nd, int64_t* counter, int advance, const uint8_t* sphase = nullptr,
                const uint8_t* dphase = nullptr, int phase = -1, double* z = nullptr, const Chain* zchain = nullptr);
int launch_shadow_init(gds_ctx* ctx, const double* y, double* z, int64_t n, const Chain& ch);
int launch_fill(gds_ctx* ctx, double* p, int64_t n, double v);
int launch_cg(gds_ctx* ctx, int what, double* x, double* r, double* p, double* ap, int64_t n, double* S, double* partial,
              int mode);
int launch_reduce(gds_ctx* ctx, const double* x, int64_t n, int kind, double* scratch);
int launch_mg_spmv(gds_ctx* ctx, int tpr, int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
                   const double* x, const double* u, const double* v, const double* w, double* out, int mode);
int launch_mg_scale(gds_ctx* ctx, int64_t n, const double* w, const double* b, double* out);
int launch_pcg_final(gds_ctx* ctx, const double* partial, double* S, int mode);
int launch_probe(gds_ctx* ctx, double* const* bufs, int n_read, int n_write, int64_t n);
int launch_halo(gds_ctx* ctx, double* vec, double* buf, const int64_t* idx, int64_t n, int scatter);
