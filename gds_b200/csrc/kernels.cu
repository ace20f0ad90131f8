// sm_100a kernels of the gds stepping core.
//
// Two kernels.
//  * `lap4_kernel<UNITW>`: the node-Laplacian laws (value = alpha*Z(L0 v) + beta*b + gamma: heat, wave, coupled-map
//    lattices, diffusion terms).  One thread owns FOUR consecutive rows of a sliced-ELL slice, so every index load is
//    one 128-bit ld.global.nc.v4.s32 and every weight / state / accumulator access a 128-bit v2.f64; the four rows
//    give each thread 4x the loads in flight of a thread-per-row kernel, which is what a latency-bound gather needs.
//  * `pass_kernel<DOMAIN, UNITW>`: everything else.  Thread-per-row over the nodes / edges / faces of the lowered
//    graph.  Every thread runs the pass's row program -- a short, warp-uniform stack code kept in the kernel
// parameter (constant) bank -- whose gather opcodes walk the sliced-ELL incidence arrays (one slice = one
// warp, so entry j of the 32 rows of a warp is a single 128 B / 256 B line) and whose result is consumed by
// the fused epilogue: Dirichlet zeroing of dy/dt + Runge-Kutta stage combine, so that each stage of each
// field is ONE pass over HBM.  fp64 throughout, no tensor cores (sparse gather), no atomics (pull form:
// fixed summation order => bitwise run-to-run reproducibility).
#include <cuda_runtime.h>
#include <math.h>

#include "internal.h"

#define BLOCK 256

// streaming (read-once) loads: bypass L1 allocation so L1 stays available for the gathered vectors
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_gather(const double* p) { return __ldg(p); }
// coherent streaming load (the accumulator is read and rewritten in place by the same thread)
__device__ __forceinline__ double ld_plain_f64(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ double sgn(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }

// ---------------------------------------------------------------------------------------------------
// gathers over the node incidence (rows = nodes)
// ---------------------------------------------------------------------------------------------------
struct SliceRange {
  int64_t base;  // first slot of this lane
  uint32_t width;
};
__device__ __forceinline__ SliceRange slice_of(const uint32_t* soff, int64_t row, int lane) {
  uint32_t o0 = __ldg(soff + (row >> 5)), o1 = __ldg(soff + (row >> 5) + 1);
  SliceRange s;
  s.base = (int64_t)o0 * GDS_SLICE + lane;
  s.width = o1 - o0;
  return s;
}

// L0 row: sum_e w_e (v[nbr] - v[row])     (gds.py:140: L0 = -B B^T)
template <bool UNITW>
__device__ __forceinline__ double n_lap(const GraphDev& g, const double* __restrict__ v, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  const double c = ld_gather(v + row);
  double acc = 0.0;
  const int32_t* nb = g.n_nbr + s.base;
  const double* w = g.n_w + s.base;
#pragma unroll 4
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t k = ld_stream_i32(nb + (int64_t)j * GDS_SLICE);
    if (UNITW) {
      acc += ld_gather(v + k) - c;  // padding entries point at the row itself
    } else {
      double wj = ld_stream_f64(w + (int64_t)j * GDS_SLICE);
      acc += wj * (ld_gather(v + k) - c);
    }
  }
  return acc;
}

// (B1 y)[row] = sum_e sigma*omega_e*y_e ; MODE 0: plain, 1: relu(+), 2: relu(-)    (gds.py:280,285,290)
template <int MODE, bool UNITW>
__device__ __forceinline__ double n_bdot(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  double acc = 0.0;
  const int32_t* ed = g.n_edge + s.base;
#pragma unroll 4
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t pk = ld_stream_i32(ed + (int64_t)j * GDS_SLICE);
    bool valid = pk >= 0;
    int32_t e = valid ? (pk >> 1) : 0;
    double sg = (pk & 1) ? 1.0 : -1.0;
    double term = UNITW ? sg * ld_gather(y + e) : sg * (__ldg(g.e_omega + e) * ld_gather(y + e));
    if (MODE == 1) term = fmax(term, 0.0);
    if (MODE == 2) term = fmax(-term, 0.0);
    acc += valid ? term : 0.0;
  }
  return acc;
}

// scalar advection by an edge field (gds.py:166-177): -sum_e sigma*w_e*v_e*y[upwind node of e]
// LIE (gds.py:179-189):                                 sum_{e: flow enters row} w_e*v_e*sigma*(y[row]-y[nbr])
template <bool LIE, bool UNITW>
__device__ __forceinline__ double n_advect(const GraphDev& g, const double* __restrict__ v,
                                           const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  const double yc = ld_gather(y + row);
  double acc = 0.0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int64_t slot = s.base + (int64_t)j * GDS_SLICE;
    int32_t pk = ld_stream_i32(g.n_edge + slot);
    int32_t nb = ld_stream_i32(g.n_nbr + slot);
    double wj = UNITW ? ((pk >= 0) ? 1.0 : 0.0) : ld_stream_f64(g.n_w + slot);  // 0 on padding
    int32_t e = (pk >= 0) ? (pk >> 1) : 0;
    double sg = (pk & 1) ? 1.0 : -1.0;
    double ve = ld_gather(v + e);
    double dir = sg * sgn(ve);
    double yn = ld_gather(y + nb);
    if (LIE) {
      acc += (dir > 0.0) ? wj * ve * sg * (yc - yn) : 0.0;
    } else {
      double src = (dir < 0.0) ? yc : yn;
      acc -= sg * wj * ve * src;
    }
  }
  return acc;
}

// upwind aggregates for edge self-advection (gds.py:326-345 with dual weights gds.py:253-262):
// d_x * { sum_in w|y|, sum_in w y^2, sum_out w|y|, sum_out w y^2 }
template <bool UNITW>
__device__ __forceinline__ double4 n_eagg(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  double p1 = 0, p2 = 0, q1 = 0, q2 = 0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int64_t slot = s.base + (int64_t)j * GDS_SLICE;
    int32_t pk = ld_stream_i32(g.n_edge + slot);
    double wj = UNITW ? ((pk >= 0) ? 1.0 : 0.0) : ld_stream_f64(g.n_w + slot);
    int32_t e = (pk >= 0) ? (pk >> 1) : 0;
    double sg = (pk & 1) ? 1.0 : -1.0;
    double ye = ld_gather(y + e);
    double dir = sg * sgn(ye);
    double a1 = wj * fabs(ye), a2 = wj * ye * ye;
    p1 += (dir > 0.0) ? a1 : 0.0;
    p2 += (dir > 0.0) ? a2 : 0.0;
    q1 += (dir < 0.0) ? a1 : 0.0;
    q2 += (dir < 0.0) ? a2 : 0.0;
  }
  double d = __ldg(g.n_dinv + row);
  return make_double4(d * p1, d * p2, d * q1, d * q2);
}

// ---------------------------------------------------------------------------------------------------
// gathers with rows = edges
// ---------------------------------------------------------------------------------------------------
template <bool UNITW>
__device__ __forceinline__ double e_gradt(const GraphDev& g, const double* __restrict__ v, int64_t row) {
  int32_t u = ld_stream_i32(g.e_u + row), w = ld_stream_i32(g.e_v + row);
  double d = ld_gather(v + w) - ld_gather(v + u);
  return UNITW ? d : ld_stream_f64(g.e_omega + row) * d;  // gds.py:154
}

template <bool UNITW>
__device__ __forceinline__ double e_curlt(const GraphDev& g, const double* __restrict__ c, int64_t row, int lane) {
  SliceRange s = slice_of(g.ef_soff, row, lane);
  double acc = 0.0;
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t pk = ld_stream_i32(g.ef_face + s.base + (int64_t)j * GDS_SLICE);
    bool valid = pk >= 0;
    int32_t f = valid ? (pk >> 1) : 0;
    double cf = ld_gather(c + f);
    acc += valid ? ((pk & 1) ? -cf : cf) : 0.0;
  }
  return UNITW ? acc : ld_stream_f64(g.e_omega + row) * acc;
}

template <bool UNITW>
__device__ __forceinline__ double e_advfin(const GraphDev& g, const double* __restrict__ y,
                                           const double* __restrict__ agg, int64_t row) {
  double ye = ld_gather(y + row);
  double s = sgn(ye);
  int32_t u = ld_stream_i32(g.e_u + row), v = ld_stream_i32(g.e_v + row);
  int32_t t = (s > 0.0) ? u : v, h = (s > 0.0) ? v : u;
  // tail wants (P1,P2) = first half of its record, head wants (Q1,Q2) = second half: one 16 B load each
  const double2 at = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)t);
  const double2 ah = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)h + 1);
  double we = UNITW ? 1.0 : ld_stream_f64(g.e_w + row);
  double s1 = we * (at.x - ah.x), s2 = we * (at.y - ah.y);
  return (s == 0.0) ? 0.0 : -(ye * s1 + s * s2);
}

// rows = faces
template <bool UNITW>
__device__ __forceinline__ double f_curl(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.fe_soff, row, lane);
  double acc = 0.0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t pk = ld_stream_i32(g.fe_edge + s.base + (int64_t)j * GDS_SLICE);
    bool valid = pk >= 0;
    int32_t e = valid ? (pk >> 1) : 0;
    double term = UNITW ? ld_gather(y + e) : __ldg(g.e_omega + e) * ld_gather(y + e);
    acc += valid ? ((pk & 1) ? -term : term) : 0.0;
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------
// the row-program interpreter
// ---------------------------------------------------------------------------------------------------
#define PUSH(...)                     \
  do {                                \
    double nv__ = (__VA_ARGS__);      \
    s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = nv__; \
  } while (0)
#define DROP1() \
  do { s0 = s1; s1 = s2; s2 = s3; s3 = s4; s4 = s5; } while (0)
#define BIN(expr)                      \
  do {                                 \
    double b = s0, a = s1;             \
    s0 = (expr);                       \
    s1 = s2; s2 = s3; s3 = s4; s4 = s5; \
  } while (0)

__device__ __forceinline__ double powi(double x, int n) {
  bool neg = n < 0;
  unsigned m = neg ? (unsigned)(-n) : (unsigned)n;
  double r = 1.0;
  while (m) {
    if (m & 1u) r *= x;
    x *= x;
    m >>= 1;
  }
  return neg ? 1.0 / r : r;
}

template <int DOMAIN, bool UNITW>
__global__ void __launch_bounds__(BLOCK) pass_kernel(const __grid_constant__ PassParams p) {
  const int64_t row = p.row0 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= p.n_rows) return;
  const int lane = threadIdx.x & 31;
  const GraphDev& g = p.g;
  // operands of the fused epilogue are requested before the row program runs, so that their latency overlaps the
  // gather chain instead of following it
  double e_yn = 0.0, e_acc = 0.0;
  uint32_t e_dbits = 0u;
  if (p.epilogue == GDS_EPI_RHS) {
    if (p.stage_mode != 4) e_yn = ld_stream_f64(p.yn + row);
    if (p.stage_mode == 1 || p.stage_mode == 2) e_acc = ld_plain_f64(p.acc_in + row);
    if (p.dmask) e_dbits = __ldg(p.dmask + (row >> 5));
  }
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
  double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  double tn, h;
  if (p.steptab) {
    StepRec rec = p.steptab[*p.step_counter];
    tn = rec.t; h = rec.h;
  } else {
    tn = p.t_fixed; h = 0.0;
  }
  for (int pc = 0; pc < p.n_code; ++pc) {
    const uint32_t ins = p.code[pc];
    const uint32_t op = ins & 0xffu, a = (ins >> 8) & 0xffu, b = (ins >> 16) & 0xffu;
    switch (op) {
      case GDS_OP_PUSH_VEC: PUSH(ld_stream_f64(p.vec[a] + row)); break;
      case GDS_OP_PUSH_IMM: PUSH(p.imm[a]); break;
      case GDS_OP_PUSH_T: PUSH(tn + p.c_stage * h); break;
      case GDS_OP_PUSH_H: PUSH(h); break;
      case GDS_OP_DUP: PUSH(s0); break;
      case GDS_OP_SWAP: { double tmp = s0; s0 = s1; s1 = tmp; } break;
      case GDS_OP_POP: DROP1(); break;
      case GDS_OP_SETL:
        if (a == 0) l0 = s0; else if (a == 1) l1 = s0; else if (a == 2) l2 = s0; else l3 = s0;
        break;
      case GDS_OP_GETL: PUSH(a == 0 ? l0 : (a == 1 ? l1 : (a == 2 ? l2 : l3))); break;
      case GDS_OP_STORE: p.out[a][row] = s0; break;
      case GDS_OP_MASK_ZERO: {
        uint32_t wbits = __ldg(p.mask[a] + (row >> 5));
        if ((wbits >> (row & 31)) & 1u) s0 = 0.0;
      } break;
      case GDS_OP_ADD: BIN(a + b); break;
      case GDS_OP_SUB: BIN(a - b); break;
      case GDS_OP_MUL: BIN(a * b); break;
      case GDS_OP_DIV: BIN(a / b); break;
      case GDS_OP_POW: BIN(pow(a, b)); break;
      case GDS_OP_MAX: BIN(fmax(a, b)); break;
      case GDS_OP_MIN: BIN(fmin(a, b)); break;
      case GDS_OP_ATAN2: BIN(atan2(a, b)); break;
      case GDS_OP_NEG: s0 = -s0; break;
      case GDS_OP_ABS: s0 = fabs(s0); break;
      case GDS_OP_SIGN: s0 = sgn(s0); break;
      case GDS_OP_SQRT: s0 = sqrt(s0); break;
      case GDS_OP_EXP: s0 = exp(s0); break;
      case GDS_OP_LOG: s0 = log(s0); break;
      case GDS_OP_SIN: s0 = sin(s0); break;
      case GDS_OP_COS: s0 = cos(s0); break;
      case GDS_OP_TAN: s0 = tan(s0); break;
      case GDS_OP_TANH: s0 = tanh(s0); break;
      case GDS_OP_SINH: s0 = sinh(s0); break;
      case GDS_OP_COSH: s0 = cosh(s0); break;
      case GDS_OP_ATAN: s0 = atan(s0); break;
      case GDS_OP_SQUARE: s0 = s0 * s0; break;
      case GDS_OP_RECIP: s0 = 1.0 / s0; break;
      case GDS_OP_POWI: s0 = powi(s0, (int)(int8_t)a); break;
      default:
        if (DOMAIN == GDS_NODES) {
          switch (op) {
            case GDS_OP_N_LAP: PUSH(n_lap<UNITW>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_BDOT: PUSH(n_bdot<0, UNITW>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_INFLUX: PUSH(n_bdot<1, UNITW>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_OUTFLUX: PUSH(n_bdot<2, UNITW>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_ADVECT: PUSH(n_advect<false, UNITW>(g, p.vec[a], p.vec[b], row, lane)); break;
            case GDS_OP_N_LIE: PUSH(n_advect<true, UNITW>(g, p.vec[a], p.vec[b], row, lane)); break;
            case GDS_OP_N_EAGG:
            {
              double4 q = n_eagg<UNITW>(g, p.vec[a], row, lane);
              double2* dst = reinterpret_cast<double2*>(p.out[b]) + 2 * row;
              dst[0] = make_double2(q.x, q.y);
              dst[1] = make_double2(q.z, q.w);
            } break;
            default: break;
          }
        } else if (DOMAIN == GDS_EDGES) {
          switch (op) {
            case GDS_OP_E_GRADT: PUSH(e_gradt<UNITW>(g, p.vec[a], row)); break;
            case GDS_OP_E_CURLT: PUSH(e_curlt<UNITW>(g, p.vec[a], row, lane)); break;
            case GDS_OP_E_ADVFIN: PUSH(e_advfin<UNITW>(g, p.vec[a], p.vec[b], row)); break;
            default: break;
          }
        } else {
          if (op == GDS_OP_F_CURL) PUSH(f_curl<UNITW>(g, p.vec[a], row, lane));
        }
        break;
    }
  }
  if (p.epilogue == GDS_EPI_RHS) {
    double k = s0;
    if ((e_dbits >> (row & 31)) & 1u) k = 0.0;  // fds.py:303
    const double yn = e_yn;
    switch (p.stage_mode) {
      case 0:  // first stage
        p.acc_out[row] = k;
        p.xnext[row] = yn + (p.a_next * h) * k;
        break;
      case 1:  // middle stage
        p.acc_out[row] = e_acc + p.acc_w * k;
        p.xnext[row] = yn + (p.a_next * h) * k;
        break;
      case 2:  // last stage
        p.ynew[row] = yn + (p.fin * h) * (e_acc + k);
        break;
      case 3:  // single stage (Euler)
        p.ynew[row] = yn + h * k;
        break;
      default:  // map: y <- map_fun(y)  (fds.py:337)
        p.ynew[row] = k;
        break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// the specialised node-Laplacian kernel: 4 consecutive rows per thread, 128-bit loads and stores
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_stream_v4i32(const int32_t* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_v2f64(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_plain_v2f64(const double* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_v2f64(double* p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
// store 4 values of a quad; `full` = all four rows exist
__device__ __forceinline__ void st_quad(double* p, int64_t r0, int64_t n_rows, bool full, const double (&x)[4]) {
  if (full) {
    st_v2f64(p + r0, x[0], x[1]);
    st_v2f64(p + r0 + 2, x[2], x[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < n_rows) p[r0 + i] = x[i];
  }
}

#define LAP4_BLOCK 128

template <bool UNITW>
__global__ void __launch_bounds__(LAP4_BLOCK) lap4_kernel(const __grid_constant__ PassParams p) {
  const int64_t r0 = p.row0 + ((int64_t)blockIdx.x * LAP4_BLOCK + threadIdx.x) * 4;
  if (r0 >= p.n_rows) return;
  const bool full = r0 + 3 < p.n_rows;
  const GraphDev& g = p.g;
  const double* __restrict__ v = p.lap_v;
  // the four rows sit in one slice (32 % 4 == 0): entry j of rows r0..r0+3 is 16 contiguous bytes of indices
  const uint32_t o0 = __ldg(g.n_soff + (r0 >> 5)), o1 = __ldg(g.n_soff + (r0 >> 5) + 1);
  const int64_t base = (int64_t)o0 * GDS_SLICE + (r0 & 31);
  const uint32_t width = o1 - o0;
  // everything that does not depend on the gather is requested first (all buffers carry >= 4 doubles of slack, so
  // the 128-bit loads of a partial last quad stay inside their allocations)
  double yn[4] = {0, 0, 0, 0}, ac[4] = {0, 0, 0, 0}, bb[4] = {0, 0, 0, 0};
  uint32_t dbits = 0u, mbits = 0u;
  const int mode = p.stage_mode;
  if (p.epilogue == GDS_EPI_RHS) {
    if (mode != 4) {
      double2 a = ld_stream_v2f64(p.yn + r0), b = ld_stream_v2f64(p.yn + r0 + 2);
      yn[0] = a.x; yn[1] = a.y; yn[2] = b.x; yn[3] = b.y;
    }
    if (mode == 1 || mode == 2) {
      double2 a = ld_plain_v2f64(p.acc_in + r0), b = ld_plain_v2f64(p.acc_in + r0 + 2);
      ac[0] = a.x; ac[1] = a.y; ac[2] = b.x; ac[3] = b.y;
    }
    if (p.dmask) dbits = __ldg(p.dmask + (r0 >> 5)) >> (r0 & 31);
  }
  if (p.lap_mask) mbits = __ldg(p.lap_mask + (r0 >> 5)) >> (r0 & 31);
  if (p.lap_b) {
    double2 a = ld_stream_v2f64(p.lap_b + r0), b = ld_stream_v2f64(p.lap_b + r0 + 2);
    bb[0] = a.x; bb[1] = a.y; bb[2] = b.x; bb[3] = b.y;
  }
  double c[4];
  {
    double2 a = __ldg(reinterpret_cast<const double2*>(v + r0)), b = __ldg(reinterpret_cast<const double2*>(v + r0 + 2));
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
  }
  double s[4] = {0, 0, 0, 0};
  const int32_t* nb = g.n_nbr + base;
  const double* w = g.n_w + base;
#pragma unroll 2
  for (uint32_t j = 0; j < width; ++j) {
    const int4 k = ld_stream_v4i32(nb + (int64_t)j * GDS_SLICE);
    if (UNITW) {
      s[0] += ld_gather(v + k.x) - c[0];
      s[1] += ld_gather(v + k.y) - c[1];
      s[2] += ld_gather(v + k.z) - c[2];
      s[3] += ld_gather(v + k.w) - c[3];
    } else {
      const double2 w01 = ld_stream_v2f64(w + (int64_t)j * GDS_SLICE), w23 = ld_stream_v2f64(w + (int64_t)j * GDS_SLICE + 2);
      s[0] += w01.x * (ld_gather(v + k.x) - c[0]);
      s[1] += w01.y * (ld_gather(v + k.y) - c[1]);
      s[2] += w23.x * (ld_gather(v + k.z) - c[2]);
      s[3] += w23.y * (ld_gather(v + k.w) - c[3]);
    }
  }
  double k4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double l = ((mbits >> i) & 1u) ? 0.0 : s[i];        // Z_mask: Dirichlet rows of L0 (gds.py:116-118)
    double val = p.lap_alpha * l;
    if (p.lap_b) val += p.lap_beta * bb[i];
    val += p.lap_gamma;
    k4[i] = val;
  }
  if (p.lap_store) st_quad(p.lap_store, r0, p.n_rows, full, k4);
  if (p.epilogue != GDS_EPI_RHS) return;
  double h = 0.0;
  if (p.steptab) h = p.steptab[*p.step_counter].h;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if ((dbits >> i) & 1u) k4[i] = 0.0;                 // fds.py:303
  double o1v[4], o2v[4];
  switch (mode) {
    case 0:
#pragma unroll
      for (int i = 0; i < 4; ++i) { o1v[i] = k4[i]; o2v[i] = yn[i] + (p.a_next * h) * k4[i]; }
      st_quad(p.acc_out, r0, p.n_rows, full, o1v);
      st_quad(p.xnext, r0, p.n_rows, full, o2v);
      break;
    case 1:
#pragma unroll
      for (int i = 0; i < 4; ++i) { o1v[i] = ac[i] + p.acc_w * k4[i]; o2v[i] = yn[i] + (p.a_next * h) * k4[i]; }
      st_quad(p.acc_out, r0, p.n_rows, full, o1v);
      st_quad(p.xnext, r0, p.n_rows, full, o2v);
      break;
    case 2:
#pragma unroll
      for (int i = 0; i < 4; ++i) o1v[i] = yn[i] + (p.fin * h) * (ac[i] + k4[i]);
      st_quad(p.ynew, r0, p.n_rows, full, o1v);
      break;
    case 3:
#pragma unroll
      for (int i = 0; i < 4; ++i) o1v[i] = yn[i] + h * k4[i];
      st_quad(p.ynew, r0, p.n_rows, full, o1v);
      break;
    default:
      st_quad(p.ynew, r0, p.n_rows, full, k4);
      break;
  }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

int launch_pass(gds_ctx* ctx, const PassParams& p, int32_t domain) {
  const int64_t rows = p.n_rows - p.row0;
  if (rows <= 0) return 0;
  const bool unitw = p.unit_weights != 0;
  if (p.lap_v && domain == GDS_NODES) {
    // 128-bit accesses need every row-0 pointer on a 16-byte boundary (the engine lays fields out that way)
    bool ok = aligned16(p.lap_v) && aligned16(p.lap_b) && aligned16(p.lap_store);
    if (p.epilogue == GDS_EPI_RHS)
      ok = ok && aligned16(p.yn) && aligned16(p.acc_in) && aligned16(p.acc_out) && aligned16(p.xnext) && aligned16(p.ynew);
    if (ok) {
      int64_t quads = (rows + 3) / 4;
      int64_t blocks = (quads + LAP4_BLOCK - 1) / LAP4_BLOCK;
      if (blocks > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
      if (unitw) lap4_kernel<true><<<(unsigned)blocks, LAP4_BLOCK, 0, ctx->stream>>>(p);
      else lap4_kernel<false><<<(unsigned)blocks, LAP4_BLOCK, 0, ctx->stream>>>(p);
      GDS_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
      return 0;
    }
  }
  int64_t blocks = (rows + BLOCK - 1) / BLOCK;
  if (blocks > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
  dim3 grid((unsigned)blocks), block(BLOCK);
  if (domain == GDS_NODES) {
    if (unitw) pass_kernel<GDS_NODES, true><<<grid, block, 0, ctx->stream>>>(p);
    else pass_kernel<GDS_NODES, false><<<grid, block, 0, ctx->stream>>>(p);
  } else if (domain == GDS_EDGES) {
    if (unitw) pass_kernel<GDS_EDGES, true><<<grid, block, 0, ctx->stream>>>(p);
    else pass_kernel<GDS_EDGES, false><<<grid, block, 0, ctx->stream>>>(p);
  } else if (domain == GDS_FACES) {
    if (unitw) pass_kernel<GDS_FACES, true><<<grid, block, 0, ctx->stream>>>(p);
    else pass_kernel<GDS_FACES, false><<<grid, block, 0, ctx->stream>>>(p);
  } else {
    GDS_FAIL("unknown domain");
  }
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
// Dirichlet overwrite after a step (fds.py:289-290,362): static values, or row `step` of the BC table
__global__ void dirichlet_kernel(double* __restrict__ y, const int64_t* __restrict__ idx,
                                 const double* __restrict__ val_static, const double* __restrict__ bc_tab, int64_t n,
                                 int64_t stride, const int64_t* __restrict__ counter) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = bc_tab ? bc_tab[(*counter) * stride + i] : val_static[i];
  y[idx[i]] = v;
}

int launch_dirichlet(gds_ctx* ctx, double* y, const int64_t* idx, const double* val_static, const double* bc_tab,
                     int64_t n, int64_t stride, const int64_t* counter) {
  if (n <= 0) return 0;
  unsigned blocks = (unsigned)((n + 255) / 256);
  dirichlet_kernel<<<blocks, 256, 0, ctx->stream>>>(y, idx, val_static, bc_tab, n, stride, counter);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

__global__ void advance_counter_kernel(int64_t* c) { *c += 1; }

int launch_advance_counter(gds_ctx* ctx, int64_t* counter) {
  advance_counter_kernel<<<1, 1, 0, ctx->stream>>>(counter);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

int launch_fill(gds_ctx* ctx, double* p, int64_t n, double v) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, n, v);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// halo pack (buf[i] = vec[idx[i]]) / unpack (vec[idx[i]] = buf[i]): boundary values of a partitioned state vector
__global__ void halo_kernel(double* __restrict__ vec, double* __restrict__ buf, const int64_t* __restrict__ idx, int64_t n,
                            int scatter) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = idx[i];
  if (scatter) vec[k] = buf[i];
  else buf[i] = vec[k];
}

int launch_halo(gds_ctx* ctx, double* vec, double* buf, const int64_t* idx, int64_t n, int scatter) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  halo_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(vec, buf, idx, n, scatter);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}
