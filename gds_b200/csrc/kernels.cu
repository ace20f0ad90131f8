// sm_100a kernels of the gds stepping core.
//
// One kernel template, `pass_kernel<DOMAIN>`: thread-per-row over the nodes / edges / faces of the lowered
// graph.  Every thread runs the pass's row program -- a short, warp-uniform stack code kept in the kernel
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
__device__ __forceinline__ double n_lap(const GraphDev& g, const double* __restrict__ v, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  const double c = ld_gather(v + row);
  double acc = 0.0;
  const int32_t* nb = g.n_nbr + s.base;
  const double* w = g.n_w + s.base;
#pragma unroll 4
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t k = ld_stream_i32(nb + (int64_t)j * GDS_SLICE);
    double wj = ld_stream_f64(w + (int64_t)j * GDS_SLICE);
    acc += wj * (ld_gather(v + k) - c);
  }
  return acc;
}

// (B1 y)[row] = sum_e sigma*omega_e*y_e ; MODE 0: plain, 1: relu(+), 2: relu(-)    (gds.py:280,285,290)
template <int MODE>
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
    double term = sg * (__ldg(g.e_omega + e) * ld_gather(y + e));
    if (MODE == 1) term = fmax(term, 0.0);
    if (MODE == 2) term = fmax(-term, 0.0);
    acc += valid ? term : 0.0;
  }
  return acc;
}

// scalar advection by an edge field (gds.py:166-177): -sum_e sigma*w_e*v_e*y[upwind node of e]
// LIE (gds.py:179-189):                                 sum_{e: flow enters row} w_e*v_e*sigma*(y[row]-y[nbr])
template <bool LIE>
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
    double wj = ld_stream_f64(g.n_w + slot);  // 0 on padding
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
__device__ __forceinline__ double4 n_eagg(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  double p1 = 0, p2 = 0, q1 = 0, q2 = 0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int64_t slot = s.base + (int64_t)j * GDS_SLICE;
    int32_t pk = ld_stream_i32(g.n_edge + slot);
    double wj = ld_stream_f64(g.n_w + slot);
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
__device__ __forceinline__ double e_gradt(const GraphDev& g, const double* __restrict__ v, int64_t row) {
  int32_t u = ld_stream_i32(g.e_u + row), w = ld_stream_i32(g.e_v + row);
  return ld_stream_f64(g.e_omega + row) * (ld_gather(v + w) - ld_gather(v + u));  // gds.py:154
}

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
  return ld_stream_f64(g.e_omega + row) * acc;
}

__device__ __forceinline__ double e_advfin(const GraphDev& g, const double* __restrict__ y,
                                           const double* __restrict__ agg, int64_t row) {
  double ye = ld_gather(y + row);
  double s = sgn(ye);
  int32_t u = ld_stream_i32(g.e_u + row), v = ld_stream_i32(g.e_v + row);
  int32_t t = (s > 0.0) ? u : v, h = (s > 0.0) ? v : u;
  // tail wants (P1,P2) = first half of its record, head wants (Q1,Q2) = second half: one 16 B load each
  const double2 at = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)t);
  const double2 ah = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)h + 1);
  double we = ld_stream_f64(g.e_w + row);
  double s1 = we * (at.x - ah.x), s2 = we * (at.y - ah.y);
  return (s == 0.0) ? 0.0 : -(ye * s1 + s * s2);
}

// rows = faces
__device__ __forceinline__ double f_curl(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.fe_soff, row, lane);
  double acc = 0.0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int32_t pk = ld_stream_i32(g.fe_edge + s.base + (int64_t)j * GDS_SLICE);
    bool valid = pk >= 0;
    int32_t e = valid ? (pk >> 1) : 0;
    double term = __ldg(g.e_omega + e) * ld_gather(y + e);
    acc += valid ? ((pk & 1) ? -term : term) : 0.0;
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------
// the row-program interpreter
// ---------------------------------------------------------------------------------------------------
#define PUSH(val)                     \
  do {                                \
    double nv__ = (val);              \
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

template <int DOMAIN>
__global__ void __launch_bounds__(BLOCK) pass_kernel(const __grid_constant__ PassParams p) {
  const int64_t row = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= p.n_rows) return;
  const int lane = threadIdx.x & 31;
  const GraphDev& g = p.g;
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
            case GDS_OP_N_LAP: PUSH(n_lap(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_BDOT: PUSH(n_bdot<0>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_INFLUX: PUSH(n_bdot<1>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_OUTFLUX: PUSH(n_bdot<2>(g, p.vec[a], row, lane)); break;
            case GDS_OP_N_ADVECT: PUSH(n_advect<false>(g, p.vec[a], p.vec[b], row, lane)); break;
            case GDS_OP_N_LIE: PUSH(n_advect<true>(g, p.vec[a], p.vec[b], row, lane)); break;
            case GDS_OP_N_EAGG:
            {
              double4 q = n_eagg(g, p.vec[a], row, lane);
              double2* dst = reinterpret_cast<double2*>(p.out[b]) + 2 * row;
              dst[0] = make_double2(q.x, q.y);
              dst[1] = make_double2(q.z, q.w);
            } break;
            default: break;
          }
        } else if (DOMAIN == GDS_EDGES) {
          switch (op) {
            case GDS_OP_E_GRADT: PUSH(e_gradt(g, p.vec[a], row)); break;
            case GDS_OP_E_CURLT: PUSH(e_curlt(g, p.vec[a], row, lane)); break;
            case GDS_OP_E_ADVFIN: PUSH(e_advfin(g, p.vec[a], p.vec[b], row)); break;
            default: break;
          }
        } else {
          if (op == GDS_OP_F_CURL) PUSH(f_curl(g, p.vec[a], row, lane));
        }
        break;
    }
  }
  if (p.epilogue == GDS_EPI_RHS) {
    double k = s0;
    if (p.dmask) {
      uint32_t wbits = __ldg(p.dmask + (row >> 5));
      if ((wbits >> (row & 31)) & 1u) k = 0.0;  // fds.py:303
    }
    const double yn = ld_stream_f64(p.yn + row);
    switch (p.stage_mode) {
      case 0:  // first stage
        p.acc_out[row] = k;
        p.xnext[row] = yn + (p.a_next * h) * k;
        break;
      case 1:  // middle stage
        p.acc_out[row] = ld_plain_f64(p.acc_in + row) + p.acc_w * k;
        p.xnext[row] = yn + (p.a_next * h) * k;
        break;
      case 2:  // last stage
        p.ynew[row] = yn + (p.fin * h) * (ld_plain_f64(p.acc_in + row) + k);
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

int launch_pass(gds_ctx* ctx, const PassParams& p, int32_t domain) {
  if (p.n_rows <= 0) return 0;
  int64_t blocks = (p.n_rows + BLOCK - 1) / BLOCK;
  if (blocks > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
  dim3 grid((unsigned)blocks), block(BLOCK);
  if (domain == GDS_NODES) pass_kernel<GDS_NODES><<<grid, block, 0, ctx->stream>>>(p);
  else if (domain == GDS_EDGES) pass_kernel<GDS_EDGES><<<grid, block, 0, ctx->stream>>>(p);
  else if (domain == GDS_FACES) pass_kernel<GDS_FACES><<<grid, block, 0, ctx->stream>>>(p);
  else GDS_FAIL("unknown domain");
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
