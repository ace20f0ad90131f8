/*
 * gds_b200 -- C-ABI of the B200-native stepping core for asrvsn/gds' hot path.
 *
 * The reference (pure Python) has no FFI of its own; its boundary for this path is the Python-level
 * field / evolution / integrator API.  This header is the C-ABI a host binds in its
 * place; each entry point cites the reference code it replaces (paths relative to the reference root).
 * The Python host in gds_b200/ binds it with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure, with the message
 * available from gds_last_error() (thread-local).  The caller owns host buffers; the library owns
 * device buffers.  One host thread drives a context; all work is ordered on the context's stream.
 * All field values are float64, all indices int32 (rows per device < 2^31, edges < 2^30).
 */
#ifndef GDS_B200_H
#define GDS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDS_ABI_VERSION 1

typedef struct gds_ctx gds_ctx;
typedef struct gds_graph gds_graph;
typedef struct gds_buf gds_buf;
typedef struct gds_mask gds_mask;
typedef struct gds_system gds_system;

/* domains of a field (gds/types.py:23-27, GraphDomain) */
enum { GDS_NODES = 0, GDS_EDGES = 1, GDS_FACES = 3 };

/* fixed-step schemes (new members beside gds/types.py:47-51 Integrators; protocol gds/ode/midpoint.py:10-32) */
enum { GDS_SCHEME_EULER = 0, GDS_SCHEME_RK4 = 1, GDS_SCHEME_MAP = 2 };

/* when a pass of the traced evolution law runs */
enum { GDS_WHEN_STAGE = 0, /* every RHS evaluation (fds.dydt, gds/fds.py:292-305)               */
       GDS_WHEN_STEP = 1   /* once per step, before stage 1: sub-expressions of accepted state only */ };

/* what a pass does with the value its row program leaves on the stack */
enum { GDS_EPI_NONE = 0,  /* nothing (the program used STORE)                                       */
       GDS_EPI_RHS = 1    /* it is dy/dt of this row: Dirichlet zeroing (fds.py:303) + stage combine  */ };

/* kinds of vector operand of a pass */
enum { GDS_VEC_BUF = 0,    /* a gds_buf                                                  */
       GDS_VEC_STATE = 1,  /* accepted state y_n of the system (`obs.y`, fds.py:373-378)  */
       GDS_VEC_STAGE = 2   /* stage state handed to dydt (`y` argument, fds.py:300-302)   */ };

/* ---- row-program opcodes: one 32-bit word  op | a<<8 | b<<16 | c<<24  -------------------------------- */
enum {
  GDS_OP_END = 0,
  /* stack sources */
  GDS_OP_PUSH_VEC = 1,   /* push vec[a][row]                      */
  GDS_OP_PUSH_IMM = 2,   /* push imm[a]                           */
  GDS_OP_PUSH_T = 3,     /* push stage time t_n + c_s h           */
  GDS_OP_PUSH_H = 4,     /* push step size h                      */
  GDS_OP_DUP = 5, GDS_OP_SWAP = 6, GDS_OP_POP = 7,
  GDS_OP_SETL = 8,       /* local[a] = top (a < 4), keeps top     */
  GDS_OP_GETL = 9,       /* push local[a]                         */
  GDS_OP_STORE = 10,     /* out[a][row] = top, keeps top          */
  GDS_OP_MASK_ZERO = 11, /* top = 0 where bit `row` of mask[a]    */
  /* binary (second-from-top OP top) */
  GDS_OP_ADD = 16, GDS_OP_SUB = 17, GDS_OP_MUL = 18, GDS_OP_DIV = 19, GDS_OP_POW = 20,
  GDS_OP_MAX = 21, GDS_OP_MIN = 22, GDS_OP_ATAN2 = 23,
  /* unary */
  GDS_OP_NEG = 32, GDS_OP_ABS = 33, GDS_OP_SIGN = 34, GDS_OP_SQRT = 35, GDS_OP_EXP = 36, GDS_OP_LOG = 37,
  GDS_OP_SIN = 38, GDS_OP_COS = 39, GDS_OP_TAN = 40, GDS_OP_TANH = 41, GDS_OP_SINH = 42, GDS_OP_COSH = 43,
  GDS_OP_ATAN = 44, GDS_OP_SQUARE = 45, GDS_OP_RECIP = 46, GDS_OP_POWI = 47 /* top ** (int8)a */,
  /* gathers over the lowered incidence structure (push the row's result) */
  GDS_OP_N_LAP = 64,      /* nodes: sum_e w_e (v[nbr]-v[row])         L0 row, gds.py:140,159
                           * c > 0: applied to f(v) instead of v, f = a chain of c <= 4 pointwise steps evaluated on the fly
                           * for the row and every gathered neighbour (no intermediate vector); step k is the pair
                           * imm[b + 2k] = GDS_CHAIN_* code, imm[b + 2k + 1] = its constant                               */
  GDS_OP_N_BDOT = 65,     /* nodes: (B1 v)[row], v on edges           gds.py:280 (div = -B1 y)          */
  GDS_OP_N_INFLUX = 66,   /* nodes: sum relu(+B1 .* v)                gds.py:282-285                    */
  GDS_OP_N_OUTFLUX = 67,  /* nodes: sum relu(-B1 .* v)                gds.py:287-290                    */
  GDS_OP_N_ADVECT = 68,   /* nodes: scalar advect(v=vec a, y=vec b)   gds.py:166-177                    */
  GDS_OP_N_LIE = 69,      /* nodes: lie_advect(v=vec a, y=vec b)      gds.py:179-189                    */
  GDS_OP_N_EAGG = 70,     /* nodes: 4 upwind aggregates of edge vec a -> out[b] as double4 (no push)    */
  GDS_OP_N_EAGGT = 71,    /* nodes: tangent of those aggregates at y=vec a in direction vec c, flow signs fixed -> out[b]
                           * (Jacobian-vector product of the self-advection, gds.py:372-395)           */
  GDS_OP_E_GRADT = 80,    /* edges: (B1^T v)[row], v on nodes         gds.py:154                        */
  GDS_OP_E_CURLT = 81,    /* edges: (B2^T v)[row], v on faces         gds.py:305                        */
  GDS_OP_E_ADVFIN = 82,   /* edges: self-advection from y=vec a, aggregates vec b   gds.py:326-345      */
  GDS_OP_E_ADVS1 = 83,    /* edges: S1 factor of the self-advection, w_e (d_t P1_t - d_h Q1_h), y=vec a, aggregates vec b */
  GDS_OP_F_CURL = 96      /* faces: (B2 v)[row], v on edges           gds.py:294                        */
};

/* steps of a pointwise chain (see GDS_OP_N_LAP): x -> x*x, x*k, x+k, k-x, -x, |x|, x*x*x */
enum { GDS_CHAIN_SQUARE = 1, GDS_CHAIN_MUL = 2, GDS_CHAIN_ADD = 3, GDS_CHAIN_RSUB = 4, GDS_CHAIN_NEG = 5,
       GDS_CHAIN_ABS = 6, GDS_CHAIN_CUBE = 7 };
#define GDS_MAX_CHAIN 4

#define GDS_MAX_CODE 96
#define GDS_MAX_IMM 24
#define GDS_MAX_VEC 16
#define GDS_MAX_MASK 6
#define GDS_MAX_OUT 4

typedef struct {
  int32_t kind;      /* GDS_VEC_*                                                       */
  gds_buf* buf;      /* for GDS_VEC_BUF                                                 */
  int64_t offset;    /* element offset of the operand's row 0 (field slice / order block) */
} gds_vecref;

typedef struct {
  int32_t domain;                 /* rows of this pass: GDS_NODES / GDS_EDGES / GDS_FACES             */
  int32_t when;                   /* GDS_WHEN_*                                                       */
  int32_t n_code;  const uint32_t* code;
  int32_t n_imm;   const double* imm;
  int32_t n_vec;   const gds_vecref* vec;
  int32_t n_mask;  gds_mask* const* mask;
  int32_t n_out;   gds_buf* const* out;  const int64_t* out_offset;  /* STORE / N_EAGG targets          */
  int32_t epilogue;               /* GDS_EPI_*                                                        */
  int64_t state_offset;           /* GDS_EPI_RHS: first state-vector element these rows advance       */
  gds_mask* dirichlet;            /* GDS_EPI_RHS: rows whose derivative is forced to 0 (may be NULL)  */
} gds_pass_desc;

/* ---- library ------------------------------------------------------------------------------------------- */
const char* gds_last_error(void);
int gds_abi_version(void);

/* device >= 0: a CUDA device.  device == -1: host-only context -- operator assembly (index maps for
 * gds_graph_export_csr) and row-program validation work, every call that would touch a GPU fails; it never computes. */
int gds_ctx_create(int device, gds_ctx** out);
int gds_ctx_destroy(gds_ctx* ctx);
int gds_ctx_set_stream(gds_ctx* ctx, void* cuda_stream); /* NULL = the context's own stream */
int gds_ctx_sync(gds_ctx* ctx);
/* device-side timing of everything enqueued between the two calls (CUDA events on the context stream) */
int gds_ctx_timer_start(gds_ctx* ctx);
int gds_ctx_timer_stop(gds_ctx* ctx, double* ms_out);
/* counters since context creation: kernels launched by this library, CUDA-graph replays */
int gds_ctx_counters(gds_ctx* ctx, int64_t* kernel_launches, int64_t* graph_replays);
/* measurement aid: achieved GB/s of a gather-free kernel that reads n_read and writes n_write float64 vectors of n
 * elements (the stream mix of one RK stage) -- the practical HBM ceiling the stage kernels are compared against */
int gds_ctx_bandwidth_probe(gds_ctx* ctx, int32_t n_read, int32_t n_write, int64_t n, int32_t reps, double* gbs_out);
/* write `bytes` of device scratch (used to flush L2 between timed iterations) */
int gds_ctx_flush_l2(gds_ctx* ctx, int64_t bytes);

/* ---- device vectors ------------------------------------------------------------------------------------ */
int gds_buf_create(gds_ctx* ctx, int64_t n, gds_buf** out);
/* non-owning view of n values at a device address (e.g. a slice of a system's state) usable as a pass operand */
int gds_buf_view(gds_ctx* ctx, void* dev_ptr, int64_t n, gds_buf** out);
int gds_buf_destroy(gds_buf* b);
int gds_buf_upload(gds_buf* b, int64_t offset, const double* host, int64_t n);
int gds_buf_download(gds_buf* b, int64_t offset, double* host, int64_t n);
int gds_buf_fill(gds_buf* b, int64_t offset, int64_t n, double value);
int gds_buf_devptr(gds_buf* b, void** dev_out, int64_t* n_out);
/* scalar reduction of n values (kind 0 sum, 1 max, 2 min) on the device, fixed grid and tree (bitwise reproducible); only
 * the 8-byte result is copied back.  Derived observables -- the projections of gds/types.py:95-105 such as the energies
 * and momenta of gds/examples/lid_driven_cavity.py:95-101 -- evaluate their vector part with gds_pass_run and finish here. */
int gds_buf_reduce(gds_buf* b, int64_t offset, int64_t n, int32_t kind, double* out);

/* row sets as bitmaps (Dirichlet rows fds.py:303, `free` / `normal` arguments gds.py:296-317) */
int gds_mask_create(gds_ctx* ctx, int64_t n_rows, const int64_t* rows, int64_t n, gds_mask** out);
int gds_mask_destroy(gds_mask* m);

/* ---- operator assembly: replaces GraphObservable.__init__ / gds.__init__ / node_gds / edge_gds ctors
 *      (gds/gds.py:16-55,105-109,136-145,194-271).  Edges are given in G.edges() order with the stored
 *      orientation (u -> v); weights NULL = all 1.0; faces as a CSR list of node indices in traversal order
 *      (n_faces = 0: no faces).  Lowers once to device sliced-ELL incidence arrays. ------------------------ */
int gds_graph_create(gds_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* edge_u, const int32_t* edge_v,
                     const double* weight, int64_t n_faces, const int64_t* face_ptr, const int32_t* face_nodes,
                     gds_graph** out);
int gds_graph_destroy(gds_graph* g);
int gds_graph_sizes(gds_graph* g, int64_t* n_nodes, int64_t* n_edges, int64_t* n_faces);
/* Bit-exact index-map export for parity checks.  kind 0: B1 rows   (node -> incident edges, sign = +1 head / -1 tail),
 *                                                 kind 1: B2 rows   (face -> edges, sign as gds.py:227-236),
 *                                                 kind 2: B2 cols   (edge -> faces).
 * Call with ptr/idx/sign NULL to get sizes (n_rows, nnz); ptr has n_rows+1 entries. */
int gds_graph_export_csr(gds_graph* g, int32_t kind, int64_t* n_rows, int64_t* nnz,
                         int64_t* ptr, int32_t* idx, int32_t* sign);
/* bytes of device memory held by the lowered operators */
int gds_graph_device_bytes(gds_graph* g, int64_t* bytes);

/* ---- eager operator application: `op(y)` on host arrays (gds.py:152-345, the `y is not None` path).
 *      Runs ONE pass (epilogue must be GDS_EPI_NONE) at time t. -------------------------------------------- */
int gds_pass_run(gds_ctx* ctx, gds_graph* g, const gds_pass_desc* pass, double t);
/* the same for MATRIX operands (an operator applied to np.eye(n), gds/examples/fluid.py:350; a block of tangent vectors):
 * `passes` is a short list evaluated in order (intermediate vectors, then the result) for each of `ncols` columns; every
 * operand / output whose buffer is listed in `strided` holds one column per `stride[k]` values and is advanced by that
 * much from column to column, the others are shared.  Validated and resolved once; one call for the whole matrix. */
int gds_passes_run_cols(gds_ctx* ctx, gds_graph* g, int32_t n_pass, const gds_pass_desc* passes, double t, int64_t ncols,
                        int32_t n_strided, gds_buf* const* strided, const int64_t* stride);
/* which kernel family would run this pass: 0 row-program interpreter, 1 Laplacian kernels, 2 term-list kernel
 * (validates the program; no device needed) */
int gds_pass_kernel_kind(const gds_pass_desc* pass, int32_t* kind_out);

/* ---- conjugate gradients for A x = b, A given as a pass that reads buffer `p` and STOREs A p into buffer `ap`
 *      (x starts at 0; symmetric positive semi-definite A with b in its range, e.g. B1 B1^T = -L0).  Replaces the dense
 *      pinv / solve of the Leray projection (gds/gds.py:249,414-427) and is the kernel of a device pressure-Poisson solve
 *      (SURVEY.md 8f-1).  All vectors hold >= n values; scalars stay on the device, the residual is read back every
 *      `check_every` iterations; stops at |r| <= rtol |b| or max_iter. ------------------------------------------------ */
int gds_cg_solve(gds_ctx* ctx, gds_graph* g, const gds_pass_desc* apply, gds_buf* x, gds_buf* b, gds_buf* r, gds_buf* p,
                 gds_buf* ap, int64_t n, double rtol, int32_t max_iter, int32_t check_every, int32_t* iters_out,
                 double* relres_out);

/* ---- multigrid preconditioner for those solves.  The reference solves its pressure problem with a general convex solver
 *      per step (gds/fds.py:97-113,317-328 via examples/fluid.py:24-38) and builds the Leray projector from a dense
 *      pseudo-inverse (gds/gds.py:249); plain conjugate gradients need O(sqrt(V)) x 10 iterations on a lattice.  Here the
 *      host builds a smoothed-aggregation hierarchy of the masked graph Laplacian once per (graph, Dirichlet set)
 *      (gds_b200/multigrid.py; gds_mg_aggregate is its aggregation sweep, host-only) and hands every level over as CSR
 *      matrices: the level matrix A_l with its damped-Jacobi weights w_l = omega / diag(A_l), the prolongation P_l
 *      (n_l x n_{l+1}) and the restriction R_l = P_l^T.  The last level (n_coarse = 0) passes the dense pseudo-inverse of
 *      its matrix as `a_*` and no weights.  One application is a symmetric V-cycle (pre / post damped-Jacobi sweeps), so it
 *      is a valid preconditioner for conjugate gradients; gds_pcg_solve is gds_cg_solve with z = M^-1 r.  The operator
 *      `apply` stays the caller's pass (the V-cycle uses its own level-0 matrix), so the solution satisfies the caller's
 *      system to `rtol` whatever the quality of the hierarchy. ---- */
typedef struct gds_mg gds_mg;
int gds_mg_aggregate(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, double theta,
                     int32_t* agg_out /* n: aggregate of every row, -1 = none */, int64_t* n_agg_out);
int gds_mg_create(gds_ctx* ctx, gds_mg** out);
int gds_mg_destroy(gds_mg* mg);
int gds_mg_add_level(gds_mg* mg, int64_t n, const int64_t* a_rowptr, const int32_t* a_col, const double* a_val,
                     const double* jacobi_w, int64_t n_coarse,
                     const int64_t* p_rowptr, const int32_t* p_col, const double* p_val,
                     const int64_t* r_rowptr, const int32_t* r_col, const double* r_val);
int gds_mg_set_sweeps(gds_mg* mg, int32_t pre, int32_t post);      /* default 1, 1; equal counts keep the cycle symmetric */
int gds_mg_info(gds_mg* mg, int32_t* n_levels, int64_t* device_bytes, int64_t* launches_per_cycle);
int gds_mg_apply(gds_mg* mg, gds_buf* r, gds_buf* z);              /* z = M^-1 r: one V-cycle (asynchronous) */
int gds_pcg_solve(gds_ctx* ctx, gds_graph* g, const gds_pass_desc* apply, gds_mg* mg, gds_buf* x, gds_buf* b, gds_buf* r,
                  gds_buf* p, gds_buf* ap, gds_buf* z, int64_t n, double rtol, int32_t max_iter, int32_t check_every,
                  int32_t* iters_out, double* relres_out);

/* ---- the stepping core: replaces fds.step / step_dydt / dydt / step_map / apply_constraints and
 *      coupled_fds.step_continuous / dydt (gds/fds.py:264-305,332-369,471-507) plus the solver object
 *      (gds/ode/midpoint.py:10-32).  One system = one concatenated state vector (fds.py:439,456-460). -------- */
int gds_system_create(gds_ctx* ctx, gds_graph* g, int64_t n_state, int32_t scheme, gds_system** out);
int gds_system_destroy(gds_system* s);
int gds_system_add_pass(gds_system* s, const gds_pass_desc* pass);
/* rows advanced with dy/dt = stage[src_offset + i] (order>1 block shift, fds.py:297-298) */
int gds_system_add_shift(gds_system* s, int64_t dst_offset, int64_t src_offset, int64_t n);
/* rows that are not advanced at all (nil systems / padding): y_{n+1} = y_n */
int gds_system_add_hold(gds_system* s, int64_t offset, int64_t n);
/* Dirichlet overwrite after every step (fds.py:289-290,362): state[idx[i]] = value[i]; dynamic = values come
 * from the table passed to gds_system_step */
int gds_system_set_dirichlet(gds_system* s, const int64_t* state_idx, const double* values, int64_t n, int32_t dynamic);
/* RK4 systems whose laws read ACCEPTED state only -- `obs.y` / operators called without an argument, the idiom of
 * gds/examples/fluid.py:33-34 and wave.py:14; in the reference `obs.y` is integrator.y, which changes at step end only
 * (fds.py:377-378) -- get the same derivative at all four stages.  With on = 1 (before finalize; every pass must be free
 * of GDS_VEC_STAGE operands and of PUSH_T, fields of order <= 2) the step runs every pass ONCE and performs the four
 * stage combines in registers, operation by operation as the staged form: results are bit-identical to on = 0, the
 * step moves a quarter of the bytes, and a partitioned system exchanges its halo once per step instead of four times. */
int gds_system_set_collapsed(gds_system* s, int32_t on);
/* Carried gather operand.  A recurrence map y <- alpha L0 f(y) + beta f(y) + ... (coupled map lattices: f(y) = 1 - a y^2;
 * a pointwise chain f applied to the map's own stage vector inside GDS_OP_N_LAP) evaluates f for the row and for EVERY
 * gathered neighbour.  With on = 1 (before finalize; honoured for map systems with one such pass on an undistributed
 * graph, ignored otherwise) the system keeps Z = f(Y) beside the state: the pass gathers Z, and every kernel that writes
 * the new state -- the pass's epilogue, the Dirichlet overwrite -- writes f of it as well; Z is rebuilt from Y at the start
 * of every stepping call (the host may have written the state in between).  f is evaluated once per row by the same
 * code, so results are bit-identical. */
int gds_system_set_shadow(gds_system* s, int32_t on);
/* capture the whole step (all passes of all stages + BC overwrite) into a CUDA graph */
int gds_system_finalize(gds_system* s);
int gds_system_set_state(gds_system* s, const double* host, int64_t offset, int64_t n);
int gds_system_get_state(gds_system* s, double* host, int64_t offset, int64_t n);
int gds_system_state_devptr(gds_system* s, void** dev_out);
/* advance n_steps: step i goes from t[i] with size h[i]; bc_values is [n_steps x n_dynamic_dirichlet]
 * (the host evaluates the user's g(t_{i+1}, x) -- fds.py:350-357 -- and only the end-of-step row is ever
 * consumed, fds.py:289-290).  Replays the captured graph n_steps times; returns without synchronising. */
int gds_system_step(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc_values);
/* the same steps enqueued kernel by kernel with a CUDA-event pair around every launch (measurement aid for the
 * roofline: per-kernel device time).  class 0/1/2 = passes over nodes/edges/faces, 3 = per-step tail; arrays of 4.
 * At most 256 steps per call; synchronises. */
int gds_system_step_timed(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc_values,
                          double* class_ms, int64_t* class_launches);
/* ---- trajectory recording: replaces the per-step host copies of System.solve (gds/system.py:38-57).  Enable before
 *      finalize; then give a device buffer of k * n_state values and step with a slot per step (-1: no snapshot): the
 *      new accepted state is copied into its slot inside the captured step; the caller downloads the buffer once. ---- */
int gds_system_enable_recording(gds_system* s);
int gds_system_set_trajectory(gds_system* s, gds_buf* traj);
int gds_system_step_record(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc_values,
                           const int64_t* rec_slot);

/* ---- ensemble stepping with HOST buffers: the caller's loop `for every member: set_initial(y0=member); step(dt);
 *      read y` (gds/fds.py:137-162,264-290 once per member) as ONE call.  Every member is a whole state held in host
 *      memory, given as n_seg segments (offset, length) of the state vector; member m is uploaded, advanced by the SAME
 *      n_steps table rows (t, h, bc_values as in gds_system_step) and its new state downloaded to host_out.  Uploads,
 *      steps and downloads of consecutive members run on three streams with double-buffered device staging, so the two
 *      PCIe directions and the kernels overlap (pinned host memory needed for that; pageable memory works, serially).
 *      host_in / host_out: [n_members x n_seg] pointers.  Synchronises.  Afterwards the system holds the last member's
 *      new state.  Results are bit-identical to set_state / step / get_state member by member.  Not for a rank of a
 *      partitioned system (ghost rows would need an exchange per upload). ---- */
int gds_system_step_ensemble(gds_system* s, int64_t n_members, int32_t n_seg, const int64_t* seg_offset,
                             const int64_t* seg_n, const double* const* host_in, double* const* host_out,
                             int64_t n_steps, const double* t, const double* h, const double* bc_values);

/* ---- stage-granular stepping, for one rank of a system partitioned over several GPUs: the host runs the halo
 *      exchange between the stages, so the step is enqueued stage by stage instead of replayed as one graph.
 *      begin_steps uploads the per-step tables; then per step: run_stage(0..n_stages-1) [exchange the vector returned by
 *      stage_output after each], end_step (Dirichlet overwrite, step counter, accepted-state flip) [exchange the state].
 *      Rows [row_begin, row_end) restrict every pass of the stage (row_end < 0: to the end; row_begin % 32 == 0), so
 *      boundary rows can be computed and sent while the interior is still running. ---------------------------------- */
int gds_system_begin_steps(gds_system* s, int64_t n_steps, const double* t, const double* h, const double* bc_values);
int gds_system_run_stage(gds_system* s, int32_t stage, int64_t row_begin, int64_t row_end);
int gds_system_stage_output(gds_system* s, int32_t stage, void** dev_out);
int gds_system_end_step(gds_system* s);
/* halo lists: positions of a state vector gathered into (pack) / scattered from (unpack) a contiguous device buffer
 * that the host hands to its transport (NCCL send/recv over NVLink) */
typedef struct gds_halo gds_halo;
int gds_halo_create(gds_ctx* ctx, const int64_t* state_idx, int64_t n, gds_halo** out);
int gds_halo_destroy(gds_halo* h);
int gds_halo_pack(gds_halo* h, const void* vec_dev, void* buf_dev);
int gds_halo_unpack(gds_halo* h, const void* buf_dev, void* vec_dev);
/* ---- peer-to-peer halo over NVLink, fused into the captured step (no host, no NCCL on the data path): after every
 *      stage a push kernel stores this rank's boundary values straight into the neighbours' ghost rows (their state
 *      buffers mapped through CUDA IPC) and publishes the exchange number in their flag table; a wait kernel polls the
 *      local flags before the next stage may gather.  Protocol: every rank exports its 5 handles (Y0 Y1 XS0 XS1 FLAGS,
 *      64 bytes each; FLAGS holds 128 words), the host transport distributes them, every rank attaches its neighbours BEFORE
 *      gds_system_finalize; then gds_system_step replays the step graph as on one GPU.  src_idx: positions in this
 *      rank's state of the values sent to neighbour q (send_ptr[q] .. send_ptr[q+1]); dst_idx: where they land in the
 *      neighbour's state.  owned_rows >= 0 (owned rows first): the RHS passes stop at that row, so ghost rows are written
 *      by their owners' pushes only.  owned_rows < 0 (deep partitions: owned rows are not a prefix): the passes run over
 *      the ghost rows as well, and every push first waits for the neighbour's "my stage is done" flag, so a neighbour's
 *      own kernels never overwrite what was pushed.
 *      The wait gives up after 20 s (a dead peer, or ranks that enqueued different numbers of steps) and raises the error
 *      flag read by gds_system_p2p_status / gds_system_p2p_error; later exchanges then return at once. --------------- */
int gds_system_ipc_export(gds_system* s, uint8_t* handles /* 5 x 64 bytes */);
int gds_system_p2p_attach(gds_system* s, int32_t my_rank, int32_t n_peers, const int32_t* peer_rank,
                          const uint8_t* peer_handles /* n_peers x 5 x 64 */, const int64_t* send_ptr,
                          const int64_t* src_idx, const int64_t* dst_idx, int64_t owned_rows);
int gds_system_p2p_status(gds_system* s, int64_t* exchanges, int64_t* error);
/* the error flag alone, read from its mirror in mapped host memory WITHOUT synchronising the stream (the host looks before
 * every call).  0 = fine; k in 1..8: the wait for neighbour k-1 timed out; 101..108: the handshake with neighbour k-101.
 * After the first timeout (GDS_B200_P2P_TIMEOUT_MS, default 20000) every later exchange of the system returns at once. */
int gds_system_p2p_error(gds_system* s, int64_t* error);
/* Overlap of the exchange with the interior rows (node-only partitions, owned_rows >= 0; call between p2p_attach and
 * finalize).  Rows [0, lo_end) and [hi_begin, owned_rows) must hold every row that is sent to or reads from a neighbour;
 * per stage they run first and are pushed at once while rows [lo_end, hi_begin) run on a second stream -- both branches
 * are part of the one captured step.  lo_end, hi_begin: multiples of 128. */
int gds_system_p2p_set_overlap(gds_system* s, int64_t lo_end, int64_t hi_begin);
/* algorithmic facts for the roofline: kernels per step and device bytes held */
int gds_system_info(gds_system* s, int64_t* kernels_per_step, int64_t* device_bytes);
/* which of the optional forms the finalized system really runs in: collapsed RK4, carried gather operand, overlapped exchange */
int gds_system_modes(gds_system* s, int32_t* collapsed, int32_t* shadow, int32_t* overlap);

/* ---- native host graphs for sizes where an nx.Graph is impossible: generators that reproduce
 *      networkx's node order, G.edges() order/orientation and adjacency order bit-exactly, plus the planar face walk
 *      of gds/utils/graph/generators.py:444-511.
 *      kind: 0 grid_2d_graph(a,b)   1 grid_graph(dim=[a,b,c])   2 triangular_lattice_graph(a,b)
 *            3 hexagonal_lattice_graph(a,b)   4 random_geometric_graph(n=a, radius=r, seed=b) ---------------------- */
typedef struct gds_hostgraph gds_hostgraph;
int gds_hostgraph_lattice(int32_t kind, int64_t a, int64_t b, int64_t c, double r, int32_t with_faces, gds_hostgraph** out);
/* any positioned plane graph: edges in G.edges() order, adjacency in G.neighbors() order; runs the face walk */
int gds_hostgraph_from_graph(int64_t n_nodes, int64_t n_edges, const int32_t* edge_u, const int32_t* edge_v,
                             const int64_t* adj_ptr, const int32_t* adj_nbr, const double* pos /* [n x 2] */,
                             gds_hostgraph** out);
int gds_hostgraph_sizes(gds_hostgraph* g, int64_t* n_nodes, int64_t* n_edges, int64_t* n_faces, int64_t* face_nnz,
                        int32_t* label_dim, int32_t* has_pos);
/* copy out what is wanted (NULL = skip): labels [n x label_dim], pos [n x 2], edges, faces as CSR of node indices */
int gds_hostgraph_copy(gds_hostgraph* g, int32_t* labels, double* pos, int32_t* edge_u, int32_t* edge_v,
                       int64_t* face_ptr, int32_t* face_nodes);
int gds_hostgraph_destroy(gds_hostgraph* g);
/* lower a native host graph straight to the device (weights NULL = all 1.0) */
int gds_graph_create_from_host(gds_ctx* ctx, gds_hostgraph* g, const double* weight, gds_graph** out);

#ifdef __cplusplus
}
#endif
#endif /* GDS_B200_H */
