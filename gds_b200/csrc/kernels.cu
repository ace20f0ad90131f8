// sm_100a kernels of the gds stepping core.
//
// Kernels.
//  * `lapT_kernel<UNITW, MODE>`: the node-Laplacian laws (value = alpha*Z(L0 v) + beta*b + gamma: heat, wave,
//    coupled-map lattices, diffusion terms) with the stage epilogue fused.  A warp owns 128 rows = 4 sliced-ELL slices,
//    each lane one row per slice, so every warp access (indices, weights, state, accumulator, and the gather of a lattice
//    neighbour) is a contiguous run of full sectors and each thread keeps 4 independent rows of loads in flight.
//    `lap4_kernel<UNITW, 1>` is its generic one-row-per-thread form, used for the < 128 rows that do not fill a group.
//  * `pass_kernel<DOMAIN, UNITW>`: everything else.  Thread-per-row over the nodes / edges / faces of the lowered
//    graph.  Every thread runs the pass's row program -- a short, warp-uniform stack code kept in the kernel
// parameter (constant) bank -- whose gather opcodes walk the sliced-ELL incidence arrays (one slice = one
// warp, so entry j of the 32 rows of a warp is a single 128 B / 256 B line) and whose result is consumed by
// the fused epilogue: Dirichlet zeroing of dy/dt + Runge-Kutta stage combine, so that each stage of each
// field is ONE pass over HBM.  fp64 throughout, no tensor cores (sparse gather), no atomics (pull form:
// fixed summation order => bitwise run-to-run reproducibility).
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"

#define BLOCK 256

// streaming (read-once) loads: bypass L1 allocation so L1 stays available for the gathered vectors.
// The scalar forms are plain (non-volatile) asm: read-only data, so the compiler may issue the index loads of several
// unrolled slots ahead of the gathers that depend on them (GDS_LD_VOLATILE=volatile restores program order for A/B runs).
#ifndef GDS_LD_VOLATILE
#define GDS_LD_VOLATILE
#endif
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p) {
  int32_t v;
  asm GDS_LD_VOLATILE("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm GDS_LD_VOLATILE("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_gather(const double* p) { return __ldg(p); }
__device__ __forceinline__ int32_t ld_stream_i16(const int16_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s16 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int2 ld_stream_v2i32(const void* p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
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

// pointwise chain f(x) (GDS_CHAIN_*): the steps are warp-uniform, evaluated in the order they were traced so that the
// result is bit-identical to a separate pointwise pass
__device__ __forceinline__ double chain_apply(const Chain& ch, double x) {
  for (int i = 0; i < ch.n; ++i) {
    double t = x;
    if (ch.pre[i] == 1) t = __dmul_rn(x, x);
    else if (ch.pre[i] == 2) t = fabs(x);
    else if (ch.pre[i] == 3) t = __dmul_rn(__dmul_rn(x, x), x);   // = powi(x, 3) of the interpreter: x * (x*x)
    x = __dadd_rn(__dmul_rn(t, ch.m[i]), ch.a[i]);  // two roundings, never contracted into an FMA
  }
  return x;
}

// L0 row: sum_e w_e (f(v[nbr]) - f(v[row]))     (gds.py:140: L0 = -B B^T; f = identity unless a chain is given)
template <bool UNITW>
__device__ __forceinline__ double n_lap(const GraphDev& g, const double* __restrict__ v, int64_t row, int lane,
                                        const Chain& ch) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  if (ch.n > 0) {
    const double c = chain_apply(ch, ld_gather(v + row));
    double acc = 0.0;
    for (uint32_t j = 0; j < s.width; ++j) {
      int32_t k = ld_stream_i32(g.n_nbr + s.base + (int64_t)j * GDS_SLICE);
      double d = chain_apply(ch, ld_gather(v + k)) - c;
      acc += UNITW ? d : ld_stream_f64(g.n_w + s.base + (int64_t)j * GDS_SLICE) * d;
    }
    return acc;
  }
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

// The slot loops below run in chunks of GCH slots: all index loads of a chunk are issued, then all gathers that depend on
// them, then the values are consumed in slot order -- two memory round trips per chunk instead of two per slot, and the
// same order of additions as the plain loop (bit-identical).  Slots past the slice width are skipped (warp-uniform).
#ifndef GCH
#define GCH 4
#endif

// (B1 y)[row] = sum_e sigma*omega_e*y_e ; MODE 0: plain, 1: relu(+), 2: relu(-)    (gds.py:280,285,290)
template <int MODE, bool UNITW>
__device__ __forceinline__ double n_bdot(const GraphDev& g, const double* __restrict__ y, int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  double acc = 0.0;
  const int32_t* ed = g.n_edge + s.base;
  for (uint32_t j0 = 0; j0 < s.width; j0 += GCH) {
    int32_t pk[GCH];
    double yv[GCH], om[GCH];
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) pk[jj] = (j0 + jj < s.width) ? ld_stream_i32(ed + (int64_t)(j0 + jj) * GDS_SLICE) : -1;
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) {
      const int32_t e = (pk[jj] >= 0) ? (pk[jj] >> 1) : 0;
      yv[jj] = ld_gather(y + e);
      om[jj] = UNITW ? 1.0 : __ldg(g.e_omega + e);
    }
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj)
      if (j0 + jj < s.width) {
        const bool valid = pk[jj] >= 0;
        const double sg = (pk[jj] & 1) ? 1.0 : -1.0;
        double term = UNITW ? sg * yv[jj] : sg * (om[jj] * yv[jj]);
        if (MODE == 1) term = fmax(term, 0.0);
        if (MODE == 2) term = fmax(-term, 0.0);
        acc += valid ? term : 0.0;
      }
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
  for (uint32_t j0 = 0; j0 < s.width; j0 += GCH) {
    int32_t pk[GCH], nb[GCH];
    double wj[GCH], ve[GCH], yn[GCH];
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) {
      const bool on = j0 + jj < s.width;
      const int64_t slot = s.base + (int64_t)(j0 + jj) * GDS_SLICE;
      pk[jj] = on ? ld_stream_i32(g.n_edge + slot) : -1;
      nb[jj] = on ? ld_stream_i32(g.n_nbr + slot) : (int32_t)row;
      wj[jj] = UNITW ? ((pk[jj] >= 0) ? 1.0 : 0.0) : (on ? ld_stream_f64(g.n_w + slot) : 0.0);  // 0 on padding
    }
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) {
      ve[jj] = ld_gather(v + ((pk[jj] >= 0) ? (pk[jj] >> 1) : 0));
      yn[jj] = ld_gather(y + nb[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj)
      if (j0 + jj < s.width) {
        const double sg = (pk[jj] & 1) ? 1.0 : -1.0;
        const double dir = sg * sgn(ve[jj]);
        if (LIE) {
          acc += (dir > 0.0) ? wj[jj] * ve[jj] * sg * (yc - yn[jj]) : 0.0;
        } else {
          const double src = (dir < 0.0) ? yc : yn[jj];
          acc -= sg * wj[jj] * ve[jj] * src;
        }
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
  for (uint32_t j0 = 0; j0 < s.width; j0 += GCH) {
    int32_t pk[GCH];
    double wj[GCH], yv[GCH];
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) {
      const bool on = j0 + jj < s.width;
      const int64_t slot = s.base + (int64_t)(j0 + jj) * GDS_SLICE;
      pk[jj] = on ? ld_stream_i32(g.n_edge + slot) : -1;
      wj[jj] = UNITW ? ((pk[jj] >= 0) ? 1.0 : 0.0) : (on ? ld_stream_f64(g.n_w + slot) : 0.0);
    }
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) yv[jj] = ld_gather(y + ((pk[jj] >= 0) ? (pk[jj] >> 1) : 0));
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj)
      if (j0 + jj < s.width) {
        const double sg = (pk[jj] & 1) ? 1.0 : -1.0;
        const double ye = yv[jj];
        const double dir = sg * sgn(ye);
        const double a1 = wj[jj] * fabs(ye), a2 = wj[jj] * ye * ye;
        p1 += (dir > 0.0) ? a1 : 0.0;
        p2 += (dir > 0.0) ? a2 : 0.0;
        q1 += (dir < 0.0) ? a1 : 0.0;
        q2 += (dir < 0.0) ? a2 : 0.0;
      }
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

// Tangent of the four aggregates in direction v with the flow directions (signs of y) held fixed -- what autograd sees
// through sign() and relu() in advect_torch (gds.py:372-395): P1' = sum in*w*s_f*v_f, P2' = sum in*w*2*y_f*v_f, same for Q.
template <bool UNITW>
__device__ __forceinline__ double4 n_eaggt(const GraphDev& g, const double* __restrict__ y, const double* __restrict__ v,
                                           int64_t row, int lane) {
  SliceRange s = slice_of(g.n_soff, row, lane);
  double p1 = 0, p2 = 0, q1 = 0, q2 = 0;
#pragma unroll 2
  for (uint32_t j = 0; j < s.width; ++j) {
    int64_t slot = s.base + (int64_t)j * GDS_SLICE;
    int32_t pk = ld_stream_i32(g.n_edge + slot);
    double wj = UNITW ? ((pk >= 0) ? 1.0 : 0.0) : ld_stream_f64(g.n_w + slot);
    int32_t e = (pk >= 0) ? (pk >> 1) : 0;
    double sg = (pk & 1) ? 1.0 : -1.0;
    double ye = ld_gather(y + e), ve = ld_gather(v + e);
    double se = sgn(ye);
    double dir = sg * se;
    double a1 = wj * (se * ve), a2 = wj * (2.0 * ye * ve);
    p1 += (dir > 0.0) ? a1 : 0.0;
    p2 += (dir > 0.0) ? a2 : 0.0;
    q1 += (dir < 0.0) ? a1 : 0.0;
    q2 += (dir < 0.0) ? a2 : 0.0;
  }
  double d = __ldg(g.n_dinv + row);
  return make_double4(d * p1, d * p2, d * q1, d * q2);
}
// S1_e = w_e (d_t P1_t - d_h Q1_h): the factor of y_e in the self-advection (0 where y_e = 0)
template <bool UNITW>
__device__ __forceinline__ double e_advs1(const GraphDev& g, const double* __restrict__ y,
                                          const double* __restrict__ agg, int64_t row) {
  double ye = ld_gather(y + row);
  double s = sgn(ye);
  int32_t u = ld_stream_i32(g.e_u + row), v = ld_stream_i32(g.e_v + row);
  int32_t t = (s > 0.0) ? u : v, h = (s > 0.0) ? v : u;
  const double2 at = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)t);
  const double2 ah = __ldg(reinterpret_cast<const double2*>(agg) + 2 * (int64_t)h + 1);
  double we = UNITW ? 1.0 : ld_stream_f64(g.e_w + row);
  return (s == 0.0) ? 0.0 : we * (at.x - ah.x);
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
  for (uint32_t j0 = 0; j0 < s.width; j0 += GCH) {
    int32_t pk[GCH];
    double yv[GCH], om[GCH];
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj)
      pk[jj] = (j0 + jj < s.width) ? ld_stream_i32(g.fe_edge + s.base + (int64_t)(j0 + jj) * GDS_SLICE) : -1;
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj) {
      const int32_t e = (pk[jj] >= 0) ? (pk[jj] >> 1) : 0;
      yv[jj] = ld_gather(y + e);
      om[jj] = UNITW ? 1.0 : __ldg(g.e_omega + e);
    }
#pragma unroll
    for (int jj = 0; jj < GCH; ++jj)
      if (j0 + jj < s.width) {
        const bool valid = pk[jj] >= 0;
        const double term = UNITW ? yv[jj] : om[jj] * yv[jj];
        acc += valid ? ((pk[jj] & 1) ? -term : term) : 0.0;
      }
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------
// fused stage epilogue: Dirichlet zeroing of dy/dt (fds.py:303) + Runge-Kutta / Euler / map combine
// ---------------------------------------------------------------------------------------------------
// A law that reads accepted state only (`obs.y`, fds.py:377-378 -- SURVEY.md App. B.3) gives the same derivative k at all
// four RK4 stages, so the step is ONE pass: the four stage combines run here in registers, operation by operation as
// the four stage passes would (k1 = k2 = k3 = k4 = k), hence bit-identical to them.
__device__ __forceinline__ double rk4_collapsed(double yn, double k, double fin_h) {
  double acc = k;                 // stage 1: acc = k1
  acc = acc + 2.0 * k;            // stage 2: acc += 2 k2
  acc = acc + 2.0 * k;            // stage 3: acc += 2 k3
  return yn + fin_h * (acc + k);  // stage 4: y_{n+1} = y_n + (h/6) (acc + k4)
}
// ... and a block of an order-2 field advanced by dy/dt = stage value of the next block (fds.py:297-298), whose own
// derivative k_v is stage-independent: its stage values are v_n, v_n + (h/2) k_v (twice), v_n + h k_v.
__device__ __forceinline__ double rk4_collapsed_shift(double yn, double vn, double kv, double h, double fin_h) {
  const double v2 = vn + (0.5 * h) * kv;
  const double v4 = vn + (1.0 * h) * kv;
  double acc = vn;
  acc = acc + 2.0 * v2;
  acc = acc + 2.0 * v2;
  return yn + fin_h * (acc + v4);
}

__device__ __forceinline__ void stage_epilogue(const PassParams& p, int64_t row, double k, double yn, double acc,
                                               uint32_t dbits, double h) {
  if ((dbits >> (row & 31)) & 1u) k = 0.0;  // fds.py:303
  switch (p.stage_mode) {
    case GDS_MODE_COLLAPSED:
      if (p.store_k) p.acc_out[row] = k;
      p.ynew[row] = rk4_collapsed(yn, k, p.fin * h);
      if (p.ynew2) p.ynew2[row] = rk4_collapsed_shift(ld_stream_f64(p.yn2 + row), yn, k, h, p.fin * h);
      break;
    case GDS_MODE_COLLAPSED_SHIFT:
      p.ynew[row] = rk4_collapsed_shift(yn, __ldg(p.aux + row), k, h, p.fin * h);
      break;
    case 0:  // first stage
      p.acc_out[row] = k;
      p.xnext[row] = yn + (p.a_next * h) * k;
      break;
    case 1:  // middle stage
      p.acc_out[row] = acc + p.acc_w * k;
      p.xnext[row] = yn + (p.a_next * h) * k;
      break;
    case 2:  // last stage
      p.ynew[row] = yn + (p.fin * h) * (acc + k);
      break;
    case 3:  // single stage (Euler)
      p.ynew[row] = yn + h * k;
      break;
    default:  // map: y <- map_fun(y)  (fds.py:337)
      p.ynew[row] = k;
      break;
  }
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
            case GDS_OP_N_LAP: {
              Chain ch;
              ch.n = (int)((ins >> 24) & 0xffu);
              for (int i = 0; i < GDS_MAX_CHAIN; ++i) {
                const int code = (i < ch.n) ? (int)p.imm[b + 2 * i] : 0;
                const double k = (i < ch.n) ? p.imm[b + 2 * i + 1] : 0.0;
                ch.pre[i] = (code == GDS_CHAIN_SQUARE) ? 1 : ((code == GDS_CHAIN_ABS) ? 2 : ((code == GDS_CHAIN_CUBE) ? 3 : 0));
                ch.m[i] = (code == GDS_CHAIN_MUL) ? k : ((code == GDS_CHAIN_RSUB || code == GDS_CHAIN_NEG) ? -1.0 : 1.0);
                ch.a[i] = (code == GDS_CHAIN_ADD || code == GDS_CHAIN_RSUB) ? k : 0.0;
              }
              PUSH(n_lap<UNITW>(g, p.vec[a], row, lane, ch));
            } break;
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
            case GDS_OP_N_EAGGT:
            {
              double4 q = n_eaggt<UNITW>(g, p.vec[a], p.vec[ins >> 24], row, lane);
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
            case GDS_OP_E_ADVS1: PUSH(e_advs1<UNITW>(g, p.vec[a], p.vec[b], row)); break;
            default: break;
          }
        } else {
          if (op == GDS_OP_F_CURL) PUSH(f_curl<UNITW>(g, p.vec[a], row, lane));
        }
        break;
    }
  }
  if (p.epilogue == GDS_EPI_RHS) stage_epilogue(p, row, s0, e_yn, e_acc, e_dbits, h);
}

// ---------------------------------------------------------------------------------------------------
// the upwind aggregates of the edge self-advection (pass 1 of gds.py:326-345, and their tangent for the Jacobian-vector
// product) as their own kernel: a one-instruction row program does not need the interpreter's stack and locals (64
// registers, 44 % occupancy there)
// ---------------------------------------------------------------------------------------------------
template <bool UNITW, bool TANGENT>
__global__ void __launch_bounds__(BLOCK) eagg_kernel(const __grid_constant__ PassParams p) {
  const int64_t row = p.row0 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= p.n_rows) return;
  const int lane = threadIdx.x & 31;
  const uint32_t ins = p.code[0];
  const uint32_t a = (ins >> 8) & 0xffu, b = (ins >> 16) & 0xffu;
  const double4 q = TANGENT ? n_eaggt<UNITW>(p.g, p.vec[a], p.vec[ins >> 24], row, lane) : n_eagg<UNITW>(p.g, p.vec[a], row, lane);
  double2* dst = reinterpret_cast<double2*>(p.out[b]) + 2 * row;
  dst[0] = make_double2(q.x, q.y);
  dst[1] = make_double2(q.z, q.w);
}

// ---------------------------------------------------------------------------------------------------
// linear combinations of masked one-hop operators (TermForm): the terms are evaluated one after another and
// accumulated -- no stack machine, one warp-uniform switch per term
// ---------------------------------------------------------------------------------------------------
// Resident blocks per SM the term-list kernels are compiled for (a register cap).  Same-box A/B on B200 (profiles/README.md
// 2.3, profiles/r2/ab/): with 5 blocks (48 registers, ~100 bytes spilled) config 2 at hex 1500^2 ran 1.66 -> 1.24 ms and
// config 3 at tri 3000^2 0.84 -> 0.63 ms -- but at the BASELINE sizes (hex 2000^2, tri 4000^2), where the passes are closer
// to the DRAM limit than to the latency limit, the same builds measured 2.12 -> 2.06 ms and 1.00 -> 1.05 ms: no gain, so the
// default stays uncapped.  -DTK_MINB=5 -DET_MINB=5 -DTK_MINB_F=8 rebuilds the capped form for smaller graphs.
#if defined(TK_MINB) && !defined(TK_MINB_F)
#define TK_MINB_F TK_MINB   // face rows
#endif
#ifdef TK_MINB
#define TK_BOUNDS __launch_bounds__(BLOCK, DOMAIN == GDS_FACES ? TK_MINB_F : TK_MINB)
#else
#define TK_BOUNDS __launch_bounds__(BLOCK)     // the compiler's own choice: 56 / 40 registers (nodes / faces)
#endif
template <int DOMAIN, bool UNITW>
__global__ void TK_BOUNDS terms_kernel(const __grid_constant__ PassParams p) {
  const int64_t row = p.row0 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= p.n_rows) return;
  const int lane = threadIdx.x & 31;
  const GraphDev& g = p.g;
  double e_yn = 0.0, e_acc = 0.0;
  uint32_t e_dbits = 0u;
  if (p.epilogue == GDS_EPI_RHS) {
    if (p.stage_mode != 4) e_yn = ld_stream_f64(p.yn + row);
    if (p.stage_mode == 1 || p.stage_mode == 2) e_acc = ld_plain_f64(p.acc_in + row);
    if (p.dmask) e_dbits = __ldg(p.dmask + (row >> 5));
  }
  double val = p.lap_gamma;
  if (p.lap_b) val += p.lap_beta * chain_apply(p.lap_bchain, ld_stream_f64(p.lap_b + row));
  if (p.lap_b2) val += p.lap_beta2 * ld_stream_f64(p.lap_b2 + row);
#pragma unroll 1
  for (int k = 0; k < p.n_terms; ++k) {
    const double* a = p.term[k].a;
    const double* b = p.term[k].b;
    uint32_t mb = 0u;
    if (p.term[k].m0) mb = __ldg(p.term[k].m0 + (row >> 5));
    if (p.term[k].m1) mb |= __ldg(p.term[k].m1 + (row >> 5));
    double gv = 0.0;
    if (DOMAIN == GDS_NODES) {
      switch (p.term[k].op) {
        case GDS_OP_N_LAP: gv = n_lap<UNITW>(g, a, row, lane, p.term[k].chain); break;
        case GDS_OP_N_BDOT: gv = n_bdot<0, UNITW>(g, a, row, lane); break;
        case GDS_OP_N_INFLUX: gv = n_bdot<1, UNITW>(g, a, row, lane); break;
        case GDS_OP_N_OUTFLUX: gv = n_bdot<2, UNITW>(g, a, row, lane); break;
        case GDS_OP_N_ADVECT: gv = n_advect<false, UNITW>(g, a, b, row, lane); break;
        default: gv = n_advect<true, UNITW>(g, a, b, row, lane); break;
      }
    } else if (DOMAIN == GDS_EDGES) {
      switch (p.term[k].op) {
        case GDS_OP_E_GRADT: gv = e_gradt<UNITW>(g, a, row); break;
        case GDS_OP_E_CURLT: gv = e_curlt<UNITW>(g, a, row, lane); break;
        default: gv = e_advfin<UNITW>(g, a, b, row); break;
      }
    } else {
      gv = f_curl<UNITW>(g, a, row, lane);
    }
    if ((mb >> (row & 31)) & 1u) gv = 0.0;
    val += p.term[k].coef * gv;
  }
  if (p.lap_store) p.lap_store[row] = val;
  if (p.epilogue == GDS_EPI_RHS) {
    double h = 0.0;
    if (p.steptab) h = p.steptab[*p.step_counter].h;
    stage_epilogue(p, row, val, e_yn, e_acc, e_dbits, h);
  }
}

// ---------------------------------------------------------------------------------------------------
// term lists over EDGE rows, phased.  terms_kernel evaluates its terms one after another, each one a dependent chain
// (endpoints -> gathers; slice offset -> face indices -> gathers): with the 4 terms of the Navier-Stokes edge law a thread
// sits through ~8 memory round trips in sequence and the kernel runs at a third of the DRAM bandwidth although it moves
// the minimum number of bytes (ncu, profiles/r2_ncu_full_ns_cavity.csv).  Here the loads of ALL terms are issued level by
// level -- (0) everything addressed by the row itself: endpoints, weights, the row's own values, the slice of the
// edge -> face table; (1) everything addressed through those: endpoint gathers, upwind aggregates, face indices;
// (2) face gathers -- and only then consumed, in term order and with the arithmetic of e_gradt / e_curlt / e_advfin,
// so the result is bit-identical to terms_kernel.  Up to ET_MAX terms, edges with at most 2 faces (every planar graph).
// ---------------------------------------------------------------------------------------------------
#define ET_MAX 4
#ifdef ET_MINB           // see TK_MINB
#define ET_BOUNDS __launch_bounds__(BLOCK, ET_MINB)
#else
#define ET_BOUNDS __launch_bounds__(BLOCK)     // 64 registers
#endif
template <bool UNITW>
__global__ void ET_BOUNDS edge_terms_kernel(const __grid_constant__ PassParams p) {
  const int64_t row = p.row0 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= p.n_rows) return;
  const int lane = threadIdx.x & 31;
  const GraphDev& g = p.g;
  // ---- level 0 ---------------------------------------------------------------------------------------------------
  const int32_t u = ld_stream_i32(g.e_u + row), v = ld_stream_i32(g.e_v + row);
  double e_yn = 0.0, e_acc = 0.0;
  uint32_t e_dbits = 0u;
  if (p.epilogue == GDS_EPI_RHS) {
    if (p.stage_mode != 4) e_yn = ld_stream_f64(p.yn + row);
    if (p.stage_mode == 1 || p.stage_mode == 2) e_acc = ld_plain_f64(p.acc_in + row);
    if (p.dmask) e_dbits = __ldg(p.dmask + (row >> 5));
  }
  double bval = 0.0, b2val = 0.0;
  if (p.lap_b) bval = ld_stream_f64(p.lap_b + row);
  if (p.lap_b2) b2val = ld_stream_f64(p.lap_b2 + row);
  double omega = 1.0, we = 1.0;
  bool any_curlt = false, any_adv = false;
#pragma unroll
  for (int k = 0; k < ET_MAX; ++k)
    if (k < p.n_terms) {
      any_curlt |= p.term[k].op == GDS_OP_E_CURLT;
      any_adv |= p.term[k].op == GDS_OP_E_ADVFIN;
    }
  if (!UNITW) {
    omega = ld_stream_f64(g.e_omega + row);
    if (any_adv) we = ld_stream_f64(g.e_w + row);
  }
  uint32_t o0 = 0u, fw = 0u;
  if (any_curlt) {
    o0 = __ldg(g.ef_soff + (row >> 5));
    fw = __ldg(g.ef_soff + (row >> 5) + 1) - o0;
  }
  uint32_t mb[ET_MAX];
  double own[ET_MAX];      // E_ADVFIN: the row's own value of its operand
#pragma unroll
  for (int k = 0; k < ET_MAX; ++k) {
    mb[k] = 0u; own[k] = 0.0;
    if (k < p.n_terms) {
      if (p.term[k].m0) mb[k] = __ldg(p.term[k].m0 + (row >> 5));
      if (p.term[k].m1) mb[k] |= __ldg(p.term[k].m1 + (row >> 5));
      if (p.term[k].op == GDS_OP_E_ADVFIN) own[k] = ld_gather(p.term[k].a + row);
    }
  }
  // ---- level 1 ---------------------------------------------------------------------------------------------------
  int32_t pk0 = -1, pk1 = -1;
  if (any_curlt) {
    const int64_t base = (int64_t)o0 * GDS_SLICE + lane;
    if (fw > 0u) pk0 = ld_stream_i32(g.ef_face + base);
    if (fw > 1u) pk1 = ld_stream_i32(g.ef_face + base + GDS_SLICE);
  }
  double r0[ET_MAX], r1[ET_MAX], r2[ET_MAX], r3[ET_MAX];
#pragma unroll
  for (int k = 0; k < ET_MAX; ++k) {
    r0[k] = r1[k] = r2[k] = r3[k] = 0.0;
    if (k < p.n_terms) {
      const double* a = p.term[k].a;
      if (p.term[k].op == GDS_OP_E_GRADT) {
        r0[k] = ld_gather(a + v);
        r1[k] = ld_gather(a + u);
      } else if (p.term[k].op == GDS_OP_E_ADVFIN) {
        const double s = sgn(own[k]);
        const int32_t t = (s > 0.0) ? u : v, h = (s > 0.0) ? v : u;
        const double2 at = __ldg(reinterpret_cast<const double2*>(p.term[k].b) + 2 * (int64_t)t);
        const double2 ah = __ldg(reinterpret_cast<const double2*>(p.term[k].b) + 2 * (int64_t)h + 1);
        r0[k] = at.x; r1[k] = at.y; r2[k] = ah.x; r3[k] = ah.y;
      }
    }
  }
  // ---- level 2 ---------------------------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < ET_MAX; ++k)
    if (k < p.n_terms && p.term[k].op == GDS_OP_E_CURLT) {
      const double* c = p.term[k].a;
      r0[k] = ld_gather(c + ((pk0 >= 0) ? (pk0 >> 1) : 0));
      r1[k] = ld_gather(c + ((pk1 >= 0) ? (pk1 >> 1) : 0));
    }
  // ---- the terms, in order ------------------------------------------------------------------------------------------
  double val = p.lap_gamma;
  if (p.lap_b) val += p.lap_beta * chain_apply(p.lap_bchain, bval);
  if (p.lap_b2) val += p.lap_beta2 * b2val;
#pragma unroll
  for (int k = 0; k < ET_MAX; ++k)
    if (k < p.n_terms) {
      double gv;
      if (p.term[k].op == GDS_OP_E_GRADT) {
        const double d = r0[k] - r1[k];                               // e_gradt
        gv = UNITW ? d : omega * d;
      } else if (p.term[k].op == GDS_OP_E_CURLT) {                    // e_curlt, width <= 2
        double acc = 0.0;
        acc += (pk0 >= 0) ? ((pk0 & 1) ? -r0[k] : r0[k]) : 0.0;
        if (fw > 1u) acc += (pk1 >= 0) ? ((pk1 & 1) ? -r1[k] : r1[k]) : 0.0;
        gv = UNITW ? acc : omega * acc;
      } else {                                                         // e_advfin
        const double ye = own[k], s = sgn(ye);
        const double s1 = we * (r0[k] - r2[k]), s2 = we * (r1[k] - r3[k]);
        gv = (s == 0.0) ? 0.0 : -(ye * s1 + s * s2);
      }
      if ((mb[k] >> (row & 31)) & 1u) gv = 0.0;
      val += p.term[k].coef * gv;
    }
  if (p.lap_store) p.lap_store[row] = val;
  if (p.epilogue == GDS_EPI_RHS) {
    double h = 0.0;
    if (p.steptab) h = p.steptab[*p.step_counter].h;
    stage_epilogue(p, row, val, e_yn, e_acc, e_dbits, h);
  }
}

// ---------------------------------------------------------------------------------------------------
// the specialised node-Laplacian kernel: 4 rows per thread, one per slice of a 4-slice group, so that every warp
// access -- indices, weights, state, accumulator, and the gather of a lattice neighbour -- is one contiguous run
// ---------------------------------------------------------------------------------------------------
#define LAP4_BLOCK 128
template <bool UNITW, int LAP4_ROWS>  // rows per thread: lane + 32*i of the warp's (32*ROWS)-row group
__global__ void __launch_bounds__(LAP4_BLOCK) lap4_kernel(const __grid_constant__ PassParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t group = p.row0 + ((int64_t)blockIdx.x * (LAP4_BLOCK / 32) + (threadIdx.x >> 5)) * (32 * LAP4_ROWS);
  if (group >= p.n_rows) return;
  const GraphDev& g = p.g;
  const double* __restrict__ v = p.lap_v;
  const int mode = p.stage_mode;
  const bool rhs = p.epilogue == GDS_EPI_RHS;
  int64_t row[LAP4_ROWS], base[LAP4_ROWS];
  uint32_t width[LAP4_ROWS], wmax = 0;
  bool valid[LAP4_ROWS];
  const int64_t s0 = group >> 5;
#pragma unroll
  for (int i = 0; i < LAP4_ROWS; ++i) {
    row[i] = group + 32 * i + lane;
    valid[i] = row[i] < p.n_rows;
    // a slice whose first row exists has offsets; lanes past the last row of a partial slice have padded slots
    const bool slice_ok = group + 32 * i < p.n_rows;
    const uint32_t o0 = slice_ok ? __ldg(g.n_soff + s0 + i) : 0u, o1 = slice_ok ? __ldg(g.n_soff + s0 + i + 1) : 0u;
    base[i] = (int64_t)o0 * GDS_SLICE + lane;
    width[i] = o1 - o0;
    wmax = max(wmax, width[i]);
    if (!valid[i]) row[i] = p.n_rows - 1;  // loads stay in range; nothing is stored for this lane
  }
  // everything that does not depend on the gather is requested first
  double yn[LAP4_ROWS], ac[LAP4_ROWS], bb[LAP4_ROWS], c[LAP4_ROWS], s[LAP4_ROWS];
  uint32_t dbit[LAP4_ROWS], mbit[LAP4_ROWS];
#pragma unroll
  for (int i = 0; i < LAP4_ROWS; ++i) {
    yn[i] = (rhs && mode != 4) ? ld_stream_f64(p.yn + row[i]) : 0.0;
    ac[i] = (rhs && (mode == 1 || mode == 2)) ? ld_plain_f64(p.acc_in + row[i]) : 0.0;
    dbit[i] = (rhs && p.dmask) ? ((__ldg(p.dmask + (row[i] >> 5)) >> (row[i] & 31)) & 1u) : 0u;
    mbit[i] = p.lap_mask ? ((__ldg(p.lap_mask + (row[i] >> 5)) >> (row[i] & 31)) & 1u) : 0u;
    bb[i] = p.lap_b ? p.lap_beta * chain_apply(p.lap_bchain, ld_stream_f64(p.lap_b + row[i])) : 0.0;
    if (p.lap_b2) bb[i] += p.lap_beta2 * ld_stream_f64(p.lap_b2 + row[i]);
    c[i] = chain_apply(p.lap_chain, ld_gather(v + row[i]));
    s[i] = 0.0;
  }
#pragma unroll 2
  for (uint32_t j = 0; j < wmax; ++j) {
    int32_t k[LAP4_ROWS];
    double w[LAP4_ROWS];
#pragma unroll
    for (int i = 0; i < LAP4_ROWS; ++i) {
      const bool on = j < width[i];  // warp-uniform
      k[i] = on ? ld_stream_i32(g.n_nbr + base[i] + (int64_t)j * GDS_SLICE) : (int32_t)row[i];
      if (!UNITW) w[i] = on ? ld_stream_f64(g.n_w + base[i] + (int64_t)j * GDS_SLICE) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < LAP4_ROWS; ++i) {
      const double d = chain_apply(p.lap_chain, ld_gather(v + k[i])) - c[i];
      s[i] += UNITW ? d : w[i] * d;
    }
  }
  double h = 0.0;
  if (rhs && p.steptab) h = p.steptab[*p.step_counter].h;
#pragma unroll
  for (int i = 0; i < LAP4_ROWS; ++i) {
    if (!valid[i]) continue;
    const int64_t r = row[i];
    double val = p.lap_alpha * (mbit[i] ? 0.0 : s[i]);      // Z_mask: Dirichlet rows of L0 (gds.py:116-118)
    if (p.lap_b || p.lap_b2) val += bb[i];
    val += p.lap_gamma;
    if (p.lap_store) p.lap_store[r] = val;
    if (!rhs) continue;
    const double k1 = dbit[i] ? 0.0 : val;                  // fds.py:303
    switch (mode) {
      case 0:
        p.acc_out[r] = k1;
        p.xnext[r] = yn[i] + (p.a_next * h) * k1;
        break;
      case 1:
        p.acc_out[r] = ac[i] + p.acc_w * k1;
        p.xnext[r] = yn[i] + (p.a_next * h) * k1;
        break;
      case 2: p.ynew[r] = yn[i] + (p.fin * h) * (ac[i] + k1); break;
      case 3: p.ynew[r] = yn[i] + h * k1; break;
      case GDS_MODE_COLLAPSED:
        if (p.store_k) p.acc_out[r] = k1;
        p.ynew[r] = rk4_collapsed(yn[i], k1, p.fin * h);
        if (p.ynew2) p.ynew2[r] = rk4_collapsed_shift(ld_stream_f64(p.yn2 + r), yn[i], k1, h, p.fin * h);
        break;
      default:
        p.ynew[r] = k1;
        if (p.shadow_out) p.shadow_out[r] = chain_apply(p.shadow_chain, k1);
        break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// lapT: the production form of the node-Laplacian stage.  A warp owns a GROUP of 128 rows = 4 consecutive slices and
// each lane one row per slice (lane + 32 i), so every warp access is a contiguous run of full sectors; the stage mode
// is a template parameter and only FULL groups are handled (the launcher sends the < 128 remaining rows to the generic
// kernel), which keeps the instruction count per row at about a third of the generic kernel's -- with thousands of
// resident threads the gather is otherwise bound by instruction issue, not by HBM.
// MODE 0/1/2: first / middle / last RK4 stage, 3: Euler, 4: map, 5: whole RK4 step of a stage-independent law,
// 7: no stage epilogue (STORE only).
// ---------------------------------------------------------------------------------------------------
#define LAPT_BLOCK 128  // measured best on B200 (128 > 256 threads; register caps and deeper unrolling lose occupancy)
#define LAPT_GROUP 128

#ifndef LAPT_LATE      // build-time knobs for A/B runs, like LAPQ_*
#define LAPT_LATE 1
#endif
#ifndef LAPT_MINB
#define LAPT_MINB 8
#endif
template <bool UNITW, int MODE, bool CHAINED, bool IDX16>
__global__ void __launch_bounds__(LAPT_BLOCK, LAPT_MINB) lapT_kernel(const __grid_constant__ PassParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t group = p.row0 + ((int64_t)blockIdx.x * (LAPT_BLOCK / 32) + (threadIdx.x >> 5)) * LAPT_GROUP;
  if (group + LAPT_GROUP > p.n_rows) return;
  const GraphDev& g = p.g;
  const double* __restrict__ v = p.lap_v;
  const int64_t r = group + lane;  // this lane's row in slice i is r + 32 i
  constexpr bool RHS = MODE != 7;
  // slice offsets of the 4 slices: 16-byte aligned (group/32 is a multiple of 4)
  // (computed when every slice has one width: the index loads then depend on nothing)
  uint32_t off[5];
  if (g.n_uniw) {
#pragma unroll
    for (int i = 0; i < 5; ++i) off[i] = ((uint32_t)(group >> 5) + i) * (uint32_t)g.n_uniw;
  } else {
    const uint4 o4 = __ldg(reinterpret_cast<const uint4*>(g.n_soff + (group >> 5)));
    off[0] = o4.x; off[1] = o4.y; off[2] = o4.z; off[3] = o4.w;
    off[4] = __ldg(g.n_soff + (group >> 5) + 4);
  }
  double yn[4], ac[4], bb[4], c[4], s[4];
  uint32_t dbit[4], mbit[4], width[4];
  const int32_t* nb[4];
  const int16_t* nb16[4];
  const double* wp[4];
  uint32_t wmax = 0;
  // the stage vectors (y_n, accumulator, b, masks) are requested before the gather loop (LAPT_LATE 0) or after it
  // (LAPT_LATE 1: fewer live registers, more resident warps, one more dependent round trip)
  auto load_streams = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (RHS && MODE != 4) yn[i] = ld_stream_f64(p.yn + r + 32 * i);
      if (MODE == 1 || MODE == 2) ac[i] = ld_plain_f64(p.acc_in + r + 32 * i);
    }
    if (p.lap_b) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bb[i] = ld_stream_f64(p.lap_b + r + 32 * i);
        if (CHAINED) bb[i] = chain_apply(p.lap_bchain, bb[i]);
        bb[i] *= p.lap_beta;
      }
    }
    if (p.lap_b2) {  // second additive vector: folded into bb (beta * g(b) + beta2 * b2)
#pragma unroll
      for (int i = 0; i < 4; ++i) bb[i] = (p.lap_b ? bb[i] : 0.0) + p.lap_beta2 * ld_stream_f64(p.lap_b2 + r + 32 * i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dbit[i] = (RHS && p.dmask) ? ((__ldg(p.dmask + (group >> 5) + i) >> lane) & 1u) : 0u;
      mbit[i] = p.lap_mask ? ((__ldg(p.lap_mask + (group >> 5) + i) >> lane) & 1u) : 0u;
    }
  };
  if (!LAPT_LATE) load_streams();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = ld_gather(v + r + 32 * i);
    if (CHAINED) c[i] = chain_apply(p.lap_chain, c[i]);
    s[i] = 0.0;
    width[i] = off[i + 1] - off[i];
    wmax = max(wmax, width[i]);
    nb[i] = g.n_nbr + (int64_t)off[i] * GDS_SLICE + lane;
    if (IDX16) nb16[i] = g.n_nbr16 + (int64_t)off[i] * GDS_SLICE + lane;
    if (!UNITW) wp[i] = g.n_w + (int64_t)off[i] * GDS_SLICE + lane;
  }
  // all index loads of a chunk of CH columns are issued before the first gather that depends on them: the chain is
  // slice offsets -> indices -> gathers, once per chunk (the 2-D lattices have 4 columns: one chunk)
  constexpr int CH = 2;
  for (uint32_t j0 = 0; j0 < wmax; j0 += CH) {
    int32_t k[CH][4];
    double w[CH][4];
#pragma unroll
    for (int jj = 0; jj < CH; ++jj)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool on = j0 + jj < width[i];  // warp-uniform
        if (IDX16) {
          const int32_t d16 = on ? ld_stream_i16(nb16[i] + (j0 + jj) * GDS_SLICE) : 0;
          k[jj][i] = (d16 == -32768) ? ld_stream_i32(nb[i] + (j0 + jj) * GDS_SLICE) : (int32_t)(r + 32 * i) + d16;
        }
        else k[jj][i] = on ? ld_stream_i32(nb[i] + (j0 + jj) * GDS_SLICE) : (int32_t)(r + 32 * i);
        if (!UNITW) w[jj][i] = on ? ld_stream_f64(wp[i] + (j0 + jj) * GDS_SLICE) : 0.0;
      }
#pragma unroll
    for (int jj = 0; jj < CH; ++jj)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double gv = ld_gather(v + k[jj][i]);
        if (CHAINED) gv = chain_apply(p.lap_chain, gv);
        const double d = gv - c[i];
        s[i] += UNITW ? d : w[jj][i] * d;
      }
  }
  if (LAPT_LATE) load_streams();
  double h = 0.0;
  if ((MODE <= 3 || MODE == 5) && p.steptab) h = p.steptab[*p.step_counter].h;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t ri = r + 32 * i;
    double val = p.lap_alpha * (mbit[i] ? 0.0 : s[i]);      // Z_mask: Dirichlet rows of L0 (gds.py:116-118)
    if (p.lap_b || p.lap_b2) val += bb[i];
    val += p.lap_gamma;
    if (p.lap_store) p.lap_store[ri] = val;
    if (!RHS) continue;
    const double k1 = dbit[i] ? 0.0 : val;                  // fds.py:303
    if (MODE == 0) {
      p.acc_out[ri] = k1;
      p.xnext[ri] = yn[i] + (p.a_next * h) * k1;
    } else if (MODE == 1) {
      p.acc_out[ri] = ac[i] + p.acc_w * k1;
      p.xnext[ri] = yn[i] + (p.a_next * h) * k1;
    } else if (MODE == 2) {
      p.ynew[ri] = yn[i] + (p.fin * h) * (ac[i] + k1);
    } else if (MODE == 3) {
      p.ynew[ri] = yn[i] + h * k1;
    } else if (MODE == 5) {
      if (p.store_k) p.acc_out[ri] = k1;
      p.ynew[ri] = rk4_collapsed(yn[i], k1, p.fin * h);
      if (p.ynew2) p.ynew2[ri] = rk4_collapsed_shift(ld_stream_f64(p.yn2 + ri), yn[i], k1, h, p.fin * h);
    } else {
      p.ynew[ri] = k1;
      if (p.shadow_out) p.shadow_out[ri] = chain_apply(p.shadow_chain, k1);
    }
  }
}

template <bool UNITW, bool CHAINED, bool IDX16>
static void launch_lapT_c(const PassParams& p, int64_t groups, cudaStream_t st) {
  const unsigned blocks = (unsigned)((groups + LAPT_BLOCK / 32 - 1) / (LAPT_BLOCK / 32));
  const int mode = (p.epilogue == GDS_EPI_RHS) ? p.stage_mode : 7;
  switch (mode) {
    case 0: lapT_kernel<UNITW, 0, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    case 1: lapT_kernel<UNITW, 1, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    case 2: lapT_kernel<UNITW, 2, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    case 3: lapT_kernel<UNITW, 3, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    case 4: lapT_kernel<UNITW, 4, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    case 5: lapT_kernel<UNITW, 5, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
    default: lapT_kernel<UNITW, 7, CHAINED, IDX16><<<blocks, LAPT_BLOCK, 0, st>>>(p); break;
  }
}

template <bool UNITW>
static void launch_lapT(const PassParams& p, int64_t groups, cudaStream_t st) {
  // the pointwise-chain instantiation (L0 f(v), beta g(b)) only when a chain is present: it costs registers
  // 16-bit relative indices (IDX16) are measured slower here (2-byte loads per lane: 0.848 vs 0.730 ms per launch on the
  // weighted 7072^2 heat problem) and are only used by the quad kernel (one 64-bit load per 4 rows: 0.500 vs 0.524 ms)
  if (p.lap_chain.n > 0 || p.lap_bchain.n > 0) launch_lapT_c<UNITW, true, false>(p, groups, st);
  else launch_lapT_c<UNITW, false, false>(p, groups, st);
}

template <bool UNITW, int ROWS>
static void launch_lap(const PassParams& p, int64_t rows, cudaStream_t st) {
  int64_t groups = (rows + 32 * ROWS - 1) / (32 * ROWS);
  int64_t blocks = (groups + LAP4_BLOCK / 32 - 1) / (LAP4_BLOCK / 32);
  lap4_kernel<UNITW, ROWS><<<(unsigned)blocks, LAP4_BLOCK, 0, st>>>(p);
}

// lapq: 4 CONSECUTIVE rows per thread, every index load one ld.global.nc.v4.s32 and every vector access a v2.f64
// (needs 16-byte aligned operands); the faster form on unit-weight graphs, where no weight stream competes for registers
#define LAPQ_BLOCK 128
#ifndef LAPQ_MINB      // build-time knobs for A/B runs (make EXTRA=-DLAPQ_MINB=8 ...); the defaults are the measured best
#define LAPQ_MINB 10
#endif
#ifndef LAPQ_MINB_W    // the weighted instantiations carry 4 weights per slot more
#define LAPQ_MINB_W 8
#endif
#ifndef LAPQ_UNROLL
#define LAPQ_UNROLL 2
#endif
constexpr int kLapqUnroll = LAPQ_UNROLL;
#ifndef LAPQ_LATE      // 1: request the stage vectors (y_n, accumulator, b) after the gather loop instead of before it --
#define LAPQ_LATE 1    // fewer live registers, one more dependent round trip
#endif
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
// 256-bit accesses (sm_100: LDG.E.256 / STG.E.256), 32-byte aligned operands
__device__ __forceinline__ void ld_stream_v4f64(const double* p, double (&x)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void ld_plain_v4f64(const double* p, double (&x)[4]) {
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void ld_gather_v4f64(const double* p, double (&x)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void st_v4f64(double* p, const double (&x)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
}
// store 4 values of a quad; `full` = all four rows exist
template <bool V256>
__device__ __forceinline__ void st_quad_v(double* p, int64_t r0, int64_t n_rows, bool full, const double (&x)[4]) {
  if (full) {
    if (V256) {
      st_v4f64(p + r0, x);
    } else {
      st_v2f64(p + r0, x[0], x[1]);
      st_v2f64(p + r0 + 2, x[2], x[3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < n_rows) p[r0 + i] = x[i];
  }
}


// RUNS: (a) every streamed vector is 32-byte aligned and moves as one 256-bit access per quad; (b) a quad whose four
// neighbours in a slot are consecutive rows (a lattice direction: the four relative offsets are equal) gathers them as
// ONE access -- a 256-bit load when the run is 32-byte aligned, the row's own registers plus one scalar for the +-1
// directions, two 128-bit loads for other even offsets -- instead of four 8-byte gathers.  Same operands, same order
// of additions: results are bit-identical to the scalar gathers.  (ncu on the scalar form: the L1 data pipe, 254
// wavefronts per 128 rows, was the busiest unit at 67 %; a run costs 8 wavefronts per slot instead of ~40.)
// COLL: the collapsed RK4 epilogue (GDS_MODE_COLLAPSED) as its own instantiation, so that the staged kernels keep their
// register budget (the collapsed form holds the row's own values until the end for the fused order-2 block shift)
// EPI 2: recurrence maps with a pointwise chain on the additive vector (beta g(b), own row only) and / or a carried gather
// operand to write (shadow_out = f(new state)) -- coupled map lattices.
#ifndef LAPQ_MINB_EPI    // EPI 2 (coupled map lattice, gather served by L1 / L2: latency-bound): 12 blocks x 128 threads, 40
#define LAPQ_MINB_EPI 12 // registers, a few spilled words -- measured 0.112 -> 0.096 ms per application at n = 4e6 (10: 0.098)
#endif
template <bool UNITW, bool IDX16, bool RUNS, int EPI>
__global__ void __launch_bounds__(LAPQ_BLOCK, EPI == 2 ? LAPQ_MINB_EPI : (EPI ? 8 : (UNITW ? LAPQ_MINB : LAPQ_MINB_W))) lapq_kernel(const __grid_constant__ PassParams p) {
  constexpr bool COLL = EPI == 1;
  const int64_t r0 = p.row0 + ((int64_t)blockIdx.x * LAPQ_BLOCK + threadIdx.x) * 4;
  if (r0 >= p.n_rows) return;
  const bool full = r0 + 3 < p.n_rows;
  const GraphDev& g = p.g;
  const double* __restrict__ v = p.lap_v;
  // the four rows sit in one slice (32 % 4 == 0): entry j of rows r0..r0+3 is 16 contiguous bytes of indices
  // (computed, not loaded, when all slices have one width: the index load then depends on nothing)
  uint32_t o0, width;
  if (g.n_uniw) {
    width = (uint32_t)g.n_uniw;
    o0 = (uint32_t)(r0 >> 5) * width;
  } else {
    o0 = __ldg(g.n_soff + (r0 >> 5));
    width = __ldg(g.n_soff + (r0 >> 5) + 1) - o0;
  }
  const int64_t base = (int64_t)o0 * GDS_SLICE + (r0 & 31);
  // everything that does not depend on the gather is requested first (all buffers carry >= 4 doubles of slack, so
  // the 128-bit loads of a partial last quad stay inside their allocations)
  double yn[4] = {0, 0, 0, 0}, ac[4] = {0, 0, 0, 0}, bb[4] = {0, 0, 0, 0};
  uint32_t dbits = 0u, mbits = 0u;
  const int mode = p.stage_mode;
  auto load_streams = [&]() {
    if (p.epilogue == GDS_EPI_RHS) {
      if (mode != 4) {
        if (RUNS) {
          ld_stream_v4f64(p.yn + r0, yn);
        } else {
          double2 a = ld_stream_v2f64(p.yn + r0), b = ld_stream_v2f64(p.yn + r0 + 2);
          yn[0] = a.x; yn[1] = a.y; yn[2] = b.x; yn[3] = b.y;
        }
      }
      if (mode == 1 || mode == 2) {
        if (RUNS) {
          ld_plain_v4f64(p.acc_in + r0, ac);
        } else {
          double2 a = ld_plain_v2f64(p.acc_in + r0), b = ld_plain_v2f64(p.acc_in + r0 + 2);
          ac[0] = a.x; ac[1] = a.y; ac[2] = b.x; ac[3] = b.y;
        }
      }
      if (p.dmask) dbits = __ldg(p.dmask + (r0 >> 5)) >> (r0 & 31);
    }
    if (p.lap_mask) mbits = __ldg(p.lap_mask + (r0 >> 5)) >> (r0 & 31);
    if (p.lap_b) {
      if (RUNS) {
        ld_stream_v4f64(p.lap_b + r0, bb);
      } else {
        double2 a = ld_stream_v2f64(p.lap_b + r0), b = ld_stream_v2f64(p.lap_b + r0 + 2);
        bb[0] = a.x; bb[1] = a.y; bb[2] = b.x; bb[3] = b.y;
      }
    }
  };
  if (!LAPQ_LATE) load_streams();
  double c[4];
  if (RUNS) {
    ld_gather_v4f64(v + r0, c);
  } else {
    double2 a = __ldg(reinterpret_cast<const double2*>(v + r0)), b = __ldg(reinterpret_cast<const double2*>(v + r0 + 2));
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
  }
  double s[4] = {0, 0, 0, 0};
  const int32_t* nb = g.n_nbr + base;
  const double* w = g.n_w + base;
#pragma unroll kLapqUnroll
  for (uint32_t j = 0; j < width; ++j) {
    int4 k;
    if (IDX16) {  // 4 x int16 (neighbour - row) in one 64-bit load
      const int2 raw = ld_stream_v2i32(g.n_nbr16 + base + (int64_t)j * GDS_SLICE);
      const uint32_t ux = (uint32_t)raw.x, uy = (uint32_t)raw.y;
      const bool escape = ((ux & 0xffffu) == 0x8000u) | ((ux >> 16) == 0x8000u) | ((uy & 0xffffu) == 0x8000u) | ((uy >> 16) == 0x8000u);
      if (escape) {  // a difference that does not fit 16 bits (rare): this quad's indices come from the int32 array
        k = ld_stream_v4i32(nb + (int64_t)j * GDS_SLICE);
      } else {
        const int32_t rr = (int32_t)r0;
        k.x = rr + (int32_t)(int16_t)(raw.x & 0xffff);
        k.y = rr + 1 + (raw.x >> 16);
        k.z = rr + 2 + (int32_t)(int16_t)(raw.y & 0xffff);
        k.w = rr + 3 + (raw.y >> 16);
      }
    } else {
      k = ld_stream_v4i32(nb + (int64_t)j * GDS_SLICE);
    }
    double gv[4];
    const int32_t d = k.x - (int32_t)r0;
    if (RUNS && (k.y == k.x + 1) & (k.z == k.x + 2) & (k.w == k.x + 3)) {
      const double* q = v + k.x;
      if ((d & 3) == 0) {
        ld_gather_v4f64(q, gv);
      } else if (d == 1) {
        gv[0] = c[1]; gv[1] = c[2]; gv[2] = c[3]; gv[3] = ld_gather(q + 3);
      } else if (d == -1) {
        gv[0] = ld_gather(q); gv[1] = c[0]; gv[2] = c[1]; gv[3] = c[2];
      } else if ((d & 1) == 0) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(q)), b = __ldg(reinterpret_cast<const double2*>(q + 2));
        gv[0] = a.x; gv[1] = a.y; gv[2] = b.x; gv[3] = b.y;
      } else {
        gv[0] = ld_gather(q); gv[1] = ld_gather(q + 1); gv[2] = ld_gather(q + 2); gv[3] = ld_gather(q + 3);
      }
    } else {
      gv[0] = ld_gather(v + k.x); gv[1] = ld_gather(v + k.y); gv[2] = ld_gather(v + k.z); gv[3] = ld_gather(v + k.w);
    }
    if (UNITW) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += gv[i] - c[i];
    } else {
      double wq[4];
      if (RUNS) {
        ld_stream_v4f64(w + (int64_t)j * GDS_SLICE, wq);
      } else {
        const double2 w01 = ld_stream_v2f64(w + (int64_t)j * GDS_SLICE), w23 = ld_stream_v2f64(w + (int64_t)j * GDS_SLICE + 2);
        wq[0] = w01.x; wq[1] = w01.y; wq[2] = w23.x; wq[3] = w23.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += wq[i] * (gv[i] - c[i]);
    }
  }
  if (LAPQ_LATE) load_streams();
  double k4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double l = ((mbits >> i) & 1u) ? 0.0 : s[i];        // Z_mask: Dirichlet rows of L0 (gds.py:116-118)
    double val = p.lap_alpha * l;
    if (p.lap_b) val += p.lap_beta * (EPI == 2 ? chain_apply(p.lap_bchain, bb[i]) : bb[i]);
    val += p.lap_gamma;
    k4[i] = val;
  }
  if (p.lap_store) st_quad_v<RUNS>(p.lap_store, r0, p.n_rows, full, k4);
  if (p.epilogue != GDS_EPI_RHS) return;
  double h = 0.0;
  if (p.steptab) h = p.steptab[*p.step_counter].h;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if ((dbits >> i) & 1u) k4[i] = 0.0;                 // fds.py:303
  double o1v[4], o2v[4];
  if (COLL) {
    if (p.store_k) st_quad_v<RUNS>(p.acc_out, r0, p.n_rows, full, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) o1v[i] = rk4_collapsed(yn[i], k4[i], p.fin * h);
    st_quad_v<RUNS>(p.ynew, r0, p.n_rows, full, o1v);
    if (p.ynew2) {   // order-2 field: the value block, dy/dt = stage value of this (derivative) block, in the same pass
      double x4[4];
      if (p.yn2 == v) {   // the Laplacian's own operand (wave equation): already in registers
#pragma unroll
        for (int i = 0; i < 4; ++i) x4[i] = c[i];
      } else if (RUNS) {
        ld_stream_v4f64(p.yn2 + r0, x4);
      } else {
        const double2 a = ld_stream_v2f64(p.yn2 + r0), b = ld_stream_v2f64(p.yn2 + r0 + 2);
        x4[0] = a.x; x4[1] = a.y; x4[2] = b.x; x4[3] = b.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) o2v[i] = rk4_collapsed_shift(x4[i], yn[i], k4[i], h, p.fin * h);
      st_quad_v<RUNS>(p.ynew2, r0, p.n_rows, full, o2v);
    }
    return;
  }
  switch (mode) {
    case 0:
#pragma unroll
      for (int i = 0; i < 4; ++i) { o1v[i] = k4[i]; o2v[i] = yn[i] + (p.a_next * h) * k4[i]; }
      st_quad_v<RUNS>(p.acc_out, r0, p.n_rows, full, o1v);
      st_quad_v<RUNS>(p.xnext, r0, p.n_rows, full, o2v);
      break;
    case 1:
#pragma unroll
      for (int i = 0; i < 4; ++i) { o1v[i] = ac[i] + p.acc_w * k4[i]; o2v[i] = yn[i] + (p.a_next * h) * k4[i]; }
      st_quad_v<RUNS>(p.acc_out, r0, p.n_rows, full, o1v);
      st_quad_v<RUNS>(p.xnext, r0, p.n_rows, full, o2v);
      break;
    case 2:
#pragma unroll
      for (int i = 0; i < 4; ++i) o1v[i] = yn[i] + (p.fin * h) * (ac[i] + k4[i]);
      st_quad_v<RUNS>(p.ynew, r0, p.n_rows, full, o1v);
      break;
    case 3:
#pragma unroll
      for (int i = 0; i < 4; ++i) o1v[i] = yn[i] + h * k4[i];
      st_quad_v<RUNS>(p.ynew, r0, p.n_rows, full, o1v);
      break;
    default:   // map: y <- value; with a carried gather operand (EPI 2) also z <- f(value)
      st_quad_v<RUNS>(p.ynew, r0, p.n_rows, full, k4);
      if (EPI == 2 && p.shadow_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o1v[i] = chain_apply(p.shadow_chain, k4[i]);
        st_quad_v<RUNS>(p.shadow_out, r0, p.n_rows, full, o1v);
      }
      break;
  }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
static bool aligned32(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; }

int launch_pass(gds_ctx* ctx, const PassParams& p, int32_t domain) {
  const int64_t rows = p.n_rows - p.row0;
  if (rows <= 0) return 0;
  const bool unitw = p.unit_weights != 0;
  if (p.lap_v && domain == GDS_NODES) {
    // full 128-row groups in lapT, the remainder (< 128 rows) in the generic kernel
    if (rows / 32 > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
    PassParams q = p;
    // measured on B200, same box, ms per launch (profiles/README.md):
    //   heat RK4 on a 7072^2 grid, unit weights: quad 0.361 (scalar-gather form with early stream loads: 0.50), lapT 0.56;
    //   random weights: quad 0.606, lapT 0.674 (early stream loads, 96 registers: 0.787)
    //   wave equation on a 400^3 grid (no 16-bit indices: the plane stride does not fit): quad 0.729, lapT 0.605;
    //   693^3 (odd strides: the quad's runs are unaligned, four scalar gathers per slot): quad 0.51, lapT 0.39
    //   CML on the 10^7-node RGG (carried operand): quad 0.241, lapT 0.239
    // so the quad kernel takes the plain L0 forms of graphs that have 16-bit relative indices (2-D lattices, where one
    // 64-bit load brings the four indices of a slot) and whose runs are aligned; lapT takes everything else: other graphs,
    // pointwise chains, a second additive vector, operands that are not 16-byte aligned.
    // GDS_B200_LAP_KERNEL=quad|lapT forces one for A/B runs.
    static const char* force = getenv("GDS_B200_LAP_KERNEL");
    const bool map_mode = p.epilogue == GDS_EPI_RHS && p.stage_mode == 4;
    const bool use_quad = (force ? !strcmp(force, "quad") : (p.g.n_nbr16 != nullptr && !p.g.n_quad_odd)) && p.lap_chain.n == 0 &&
                          (p.lap_bchain.n == 0 || map_mode) && !p.lap_b2 && (!p.shadow_out || map_mode);
    bool quad_ok = use_quad && aligned16(p.lap_v) && aligned16(p.lap_b) && aligned16(p.lap_store) && (p.row0 % 4 == 0);
    if (quad_ok && p.epilogue == GDS_EPI_RHS)
      quad_ok = aligned16(p.yn) && aligned16(p.acc_in) && aligned16(p.acc_out) && aligned16(p.xnext) && aligned16(p.ynew) &&
                aligned16(p.yn2) && aligned16(p.ynew2) && aligned16(p.shadow_out);
    if (quad_ok) {
      int64_t blocks = ((rows + 3) / 4 + LAPQ_BLOCK - 1) / LAPQ_BLOCK;
      // 256-bit accesses and run gathers need 32-byte aligned vectors (the engine lays fields out that way);
      // GDS_B200_LAPQ_RUNS=0 keeps the 128-bit / scalar-gather form for A/B runs
      static const char* runs_env = getenv("GDS_B200_LAPQ_RUNS");
      bool runs = !(runs_env && !strcmp(runs_env, "0")) && aligned32(p.lap_v) && aligned32(p.lap_b) && aligned32(p.lap_store);
      if (runs && p.epilogue == GDS_EPI_RHS)
        runs = aligned32(p.yn) && aligned32(p.acc_in) && aligned32(p.acc_out) && aligned32(p.xnext) && aligned32(p.ynew) &&
               aligned32(p.yn2) && aligned32(p.ynew2) && aligned32(p.shadow_out);
      const bool coll = p.epilogue == GDS_EPI_RHS && p.stage_mode == GDS_MODE_COLLAPSED;
      const int epi = coll ? 1 : ((map_mode && (p.lap_bchain.n > 0 || p.shadow_out)) ? 2 : 0);
      const int variant = epi * 8 + ((unitw ? 4 : 0) | (p.g.n_nbr16 ? 2 : 0) | (runs ? 1 : 0));
      const unsigned nb = (unsigned)blocks;
#define LAPQ_CASE(V, A, B, C, D) case V: lapq_kernel<A, B, C, D><<<nb, LAPQ_BLOCK, 0, ctx->stream>>>(p); break;
#define LAPQ_CASES(E)                                                                                    \
        LAPQ_CASE(E * 8 + 7, true, true, true, E) LAPQ_CASE(E * 8 + 6, true, true, false, E)             \
        LAPQ_CASE(E * 8 + 5, true, false, true, E) LAPQ_CASE(E * 8 + 4, true, false, false, E)           \
        LAPQ_CASE(E * 8 + 3, false, true, true, E) LAPQ_CASE(E * 8 + 2, false, true, false, E)           \
        LAPQ_CASE(E * 8 + 1, false, false, true, E) LAPQ_CASE(E * 8 + 0, false, false, false, E)
      switch (variant) {
        LAPQ_CASES(0) LAPQ_CASES(1) LAPQ_CASES(2)
        default: GDS_FAIL("internal: no such quad kernel variant");
      }
#undef LAPQ_CASES
#undef LAPQ_CASE
      GDS_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
      return 0;
    }
    const int64_t groups = rows / LAPT_GROUP;
    if (groups > 0) {
      if (unitw) launch_lapT<true>(q, groups, ctx->stream); else launch_lapT<false>(q, groups, ctx->stream);
      GDS_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    q.row0 = p.row0 + groups * LAPT_GROUP;
    if (q.row0 < q.n_rows) {
      if (unitw) launch_lap<true, 1>(q, q.n_rows - q.row0, ctx->stream); else launch_lap<false, 1>(q, q.n_rows - q.row0, ctx->stream);
      GDS_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    return 0;
  }
  int64_t blocks = (rows + BLOCK - 1) / BLOCK;
  if (blocks > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
  dim3 grid((unsigned)blocks), block(BLOCK);
  if (p.use_terms) {
    if (domain == GDS_NODES) {
      if (unitw) terms_kernel<GDS_NODES, true><<<grid, block, 0, ctx->stream>>>(p);
      else terms_kernel<GDS_NODES, false><<<grid, block, 0, ctx->stream>>>(p);
    } else if (domain == GDS_EDGES) {
      // level-by-level loads when the list fits the phased kernel (<= 4 terms, every edge on <= 2 faces)
      const char* no_phased = getenv("GDS_B200_NO_PHASED");   // A/B and bit-identity tests; read at enqueue / capture time
      if (!no_phased && p.n_terms > 0 && p.n_terms <= ET_MAX && p.g.ef_maxw <= 2) {
        if (unitw) edge_terms_kernel<true><<<grid, block, 0, ctx->stream>>>(p);
        else edge_terms_kernel<false><<<grid, block, 0, ctx->stream>>>(p);
      } else if (unitw) terms_kernel<GDS_EDGES, true><<<grid, block, 0, ctx->stream>>>(p);
      else terms_kernel<GDS_EDGES, false><<<grid, block, 0, ctx->stream>>>(p);
    } else if (domain == GDS_FACES) {
      if (unitw) terms_kernel<GDS_FACES, true><<<grid, block, 0, ctx->stream>>>(p);
      else terms_kernel<GDS_FACES, false><<<grid, block, 0, ctx->stream>>>(p);
    } else {
      GDS_FAIL("unknown domain");
    }
    GDS_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    return 0;
  }
  if (domain == GDS_NODES && p.n_code == 1 && p.epilogue == GDS_EPI_NONE &&
      ((p.code[0] & 0xffu) == GDS_OP_N_EAGG || (p.code[0] & 0xffu) == GDS_OP_N_EAGGT)) {
    const bool tangent = (p.code[0] & 0xffu) == GDS_OP_N_EAGGT;
    if (unitw) { if (tangent) eagg_kernel<true, true><<<grid, block, 0, ctx->stream>>>(p); else eagg_kernel<true, false><<<grid, block, 0, ctx->stream>>>(p); }
    else { if (tangent) eagg_kernel<false, true><<<grid, block, 0, ctx->stream>>>(p); else eagg_kernel<false, false><<<grid, block, 0, ctx->stream>>>(p); }
  } else if (domain == GDS_NODES) {
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

// The whole per-step tail in one launch: static and time-varying Dirichlet overwrite, and (advance != 0) the step
// counter, bumped by the last block to finish so that every block has read the step's table row first.
// counter[0] = step, counter[1] = block ticket.
__global__ void __launch_bounds__(256) tail_kernel(double* __restrict__ y, const int64_t* __restrict__ sidx,
                                                   const double* __restrict__ sval, int64_t ns,
                                                   const int64_t* __restrict__ didx, const double* __restrict__ tab,
                                                   int64_t nd, int64_t* counter, int advance,
                                                   const uint8_t* __restrict__ sphase, const uint8_t* __restrict__ dphase,
                                                   int phase, double* __restrict__ z, const Chain zchain) {
  const int64_t step = counter[0];
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  // phase >= 0: only the entries of that phase (overlap split: boundary ranges first, the rest after the interior rows)
  // z: the carried gather operand f(y) follows the overwritten rows
  if (i < ns) {
    if (phase < 0 || !sphase || sphase[i] == phase) {
      y[sidx[i]] = sval[i];
      if (z) z[sidx[i]] = chain_apply(zchain, sval[i]);
    }
  } else if (i < ns + nd) {
    if (phase < 0 || !dphase || dphase[i - ns] == phase) {
      const double v = tab[step * nd + (i - ns)];
      y[didx[i - ns]] = v;
      if (z) z[didx[i - ns]] = chain_apply(zchain, v);
    }
  }
  if (!advance) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(counter + 1), 1ull);
    if (t == gridDim.x - 1) { counter[1] = 0; counter[0] = step + 1; }
  }
}

int launch_tail(gds_ctx* ctx, double* y, const int64_t* sidx, const double* sval, int64_t ns, const int64_t* didx,
                const double* tab, int64_t nd, int64_t* counter, int advance, const uint8_t* sphase,
                const uint8_t* dphase, int phase, double* z, const Chain* zchain) {
  if (ns + nd <= 0 && !advance) return 0;
  const unsigned blocks = (unsigned)std::max<int64_t>(1, (ns + nd + 255) / 256);
  Chain zc = {};
  if (z && zchain) zc = *zchain;
  tail_kernel<<<blocks, 256, 0, ctx->stream>>>(y, sidx, sval, ns, didx, tab, nd, counter, advance, sphase, dphase, phase,
                                               z, zc);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// carried gather operand: z = f(y) over the whole state (at the start of every stepping call: the host may have written y)
__global__ void __launch_bounds__(256) shadow_init_kernel(const double* __restrict__ y, double* __restrict__ z, int64_t n,
                                                          const Chain ch) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    z[i] = chain_apply(ch, ld_stream_f64(y + i));
}

int launch_shadow_init(gds_ctx* ctx, const double* y, double* z, int64_t n, const Chain& ch) {
  if (n <= 0) return 0;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 16);
  shadow_init_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(y, z, n, ch);
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

// ---------------------------------------------------------------------------------------------------
// bandwidth probe: a gather-free kernel with the stream mix of one RK stage (n_read input vectors, n_write output
// vectors, fully coalesced 128-bit accesses) -- the practical HBM ceiling the stage kernels are compared against
// ---------------------------------------------------------------------------------------------------
struct ProbeParams {
  const double* in[8];
  double* out[4];
  int n_read, n_write;
  int64_t n;
};

__global__ void __launch_bounds__(256) probe_kernel(const __grid_constant__ ProbeParams p) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
  if (i + 1 >= p.n) return;
  double2 acc = make_double2(0.0, 0.0);
  double2 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < p.n_read) {
      asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "l"(p.in[k] + i));
    }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < p.n_read) { acc.x += v[k].x; acc.y += v[k].y; }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < p.n_write) {
      asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p.out[k] + i), "d"(acc.x + k), "d"(acc.y) : "memory");
    }
}

int launch_probe(gds_ctx* ctx, double* const* bufs, int n_read, int n_write, int64_t n) {
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < n_read; ++k) p.in[k] = bufs[k];
  for (int k = 0; k < n_write; ++k) p.out[k] = bufs[n_read + k];
  p.n_read = n_read; p.n_write = n_write; p.n = n;
  int64_t blocks = (n / 2 + 255) / 256;
  probe_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// conjugate gradients: vector kernels with the scalars (alpha, beta, residual norms) resident on the device, so an
// iteration needs no host round trip.  Reductions run on a fixed grid with a fixed tree: bitwise reproducible.
// S[0] = rr, S[1] = pAp, S[2] = alpha, S[3] = beta, S[4] = bb (|b|^2), S[5] = target for rr, S[6] = finished flag
// ---------------------------------------------------------------------------------------------------
#define CG_BLOCKS 1184  // 148 SMs x 8
#define CG_THREADS 256

__device__ __forceinline__ void block_reduce_store(double v, double* out) {
  __shared__ double sh[CG_THREADS];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = CG_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

__global__ void __launch_bounds__(CG_THREADS) cg_dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                            int64_t n, double* __restrict__ partial) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS)
    acc += a[i] * b[i];
  block_reduce_store(acc, partial + blockIdx.x);
}

// mode 0: S.pAp = v, S.alpha = rr / v;   mode 1: S.beta = v / rr, S.rr = v;   mode 2 (start): S.rr = S.bb = v
__global__ void __launch_bounds__(CG_THREADS) cg_final_kernel(const double* __restrict__ partial, double* __restrict__ S,
                                                              int mode) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < CG_BLOCKS; i += CG_THREADS) acc += partial[i];
  __shared__ double total;
  block_reduce_store(acc, &total);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double v = total;
    // S[5] = convergence target for r.r, S[6] = 1 once converged or broken down: later iterations of the same batch
    // (the host only looks every `check_every` iterations) must not touch x any more -- with r at rounding level
    // p.Ap underflows faster than r.r and alpha would blow up
    if (mode == 0) {
      S[1] = v;
      const bool live = S[6] == 0.0 && v > 0.0 && isfinite(v);
      if (!live) S[6] = 1.0;
      S[2] = live ? S[0] / v : 0.0;
    } else if (mode == 1) {
      S[3] = (S[6] == 0.0 && S[0] != 0.0) ? v / S[0] : 0.0;
      S[0] = v;
      if (!(v > S[5])) S[6] = 1.0;
    } else {
      S[0] = v; S[4] = v;
    }
  }
}

// x += alpha p;  r -= alpha Ap;  partial sums of the new r.r
__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(double* __restrict__ x, double* __restrict__ r,
                                                               const double* __restrict__ p, const double* __restrict__ ap,
                                                               int64_t n, const double* __restrict__ S,
                                                               double* __restrict__ partial) {
  const double alpha = S[2];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    acc += ri * ri;
  }
  block_reduce_store(acc, partial + blockIdx.x);
}

// p = r + beta p
__global__ void __launch_bounds__(CG_THREADS) cg_direction_kernel(double* __restrict__ p, const double* __restrict__ r,
                                                                  int64_t n, const double* __restrict__ S) {
  const double beta = S[3];
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS)
    p[i] = r[i] + beta * p[i];
}

// total of the per-block partial sums, computed by EVERY block for itself in the fixed order cg_final_kernel uses (thread t
// adds partial[t], partial[t + 256], ...; then the block tree): the same bits in every block, no extra launch
__device__ __forceinline__ double cg_total(const double* __restrict__ partial) {
  __shared__ double sh[CG_THREADS];
  __shared__ double total;
  double acc = 0.0;
  for (int i = threadIdx.x; i < CG_BLOCKS; i += CG_THREADS) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = CG_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) total = sh[0];
  __syncthreads();
  const double v = total;
  __syncthreads();     // `sh` is reused by the reduction that follows in the same kernel
  return v;
}

// Fused iteration (4 launches: A p, p.Ap partials, this, the direction update).  Scalars written by one block while the
// others still read them go to shadow slots and are committed by a LATER kernel: S[7] = new r.r (committed to S[0]),
// S[8] / S[9] = "finished" raised by the update / by the direction kernel (committed to S[6]).
// p.Ap partial sums; block 0 first commits the scalars the previous direction update left in the shadow slots
__global__ void __launch_bounds__(CG_THREADS) cg_dot_commit_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                                   int64_t n, double* __restrict__ partial, double* S) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    S[0] = S[7];
    if (S[8] != 0.0 || S[9] != 0.0) S[6] = 1.0;
  }
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS)
    acc += a[i] * b[i];
  block_reduce_store(acc, partial + blockIdx.x);
}

// alpha = r.r / p.Ap from the partial sums;  x += alpha p;  r -= alpha Ap;  partial sums of the new r.r into partial2
__global__ void __launch_bounds__(CG_THREADS) cg_update_fused_kernel(double* __restrict__ x, double* __restrict__ r,
                                                                     const double* __restrict__ p, const double* __restrict__ ap,
                                                                     int64_t n, double* S, const double* __restrict__ partial,
                                                                     double* __restrict__ partial2) {
  const double v = cg_total(partial);                       // p.Ap
  const bool live = S[6] == 0.0 && v > 0.0 && isfinite(v);  // S[6], S[0] are not written in this kernel
  const double alpha = live ? S[0] / v : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    S[1] = v; S[2] = alpha;
    if (!live) S[8] = 1.0;
  }
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    acc += ri * ri;
  }
  block_reduce_store(acc, partial2 + blockIdx.x);
}

// beta = r.r_new / r.r_old from the partial sums;  p = r + beta p
__global__ void __launch_bounds__(CG_THREADS) cg_direction_fused_kernel(double* __restrict__ p, const double* __restrict__ r,
                                                                        int64_t n, double* S, const double* __restrict__ partial2) {
  const double v = cg_total(partial2);                      // new r.r
  const bool stopped = S[6] != 0.0 || S[8] != 0.0;          // S[8] was last written by the kernel before this one
  const double beta = (!stopped && S[0] != 0.0) ? v / S[0] : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    S[3] = beta; S[7] = v;
    if (!(v > S[5])) S[9] = 1.0;                            // read by nobody in this kernel
  }
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS)
    p[i] = r[i] + beta * p[i];
}

// ---------------------------------------------------------------------------------------------------
// multigrid preconditioner (gds_mg_*): every operator of the hierarchy -- level matrices, prolongations, restrictions, the
// dense pseudo-inverse of the coarsest level -- is a CSR matrix applied by ONE kernel family.  TPR threads share a row
// (lane-strided partial sums, then a fixed shuffle tree: bitwise reproducible); what happens to the row sum is the mode:
//   0: out = A x            (restriction, coarsest solve)        1: out = u - A x              (residual)
//   2: out = u + A x        (prolongation, in place: u == out)   3: out = v + w .* (u - A x)   (damped Jacobi sweep)
// ---------------------------------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256) mg_spmv_kernel(int64_t n, const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col, const double* __restrict__ val,
                                                      const double* __restrict__ x, const double* u, const double* v,
                                                      const double* __restrict__ w, double* out, int mode) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = t / TPR;
  const int lane = (int)(t % TPR);
  double acc = 0.0;
  if (row < n) {
    const int32_t k1 = rowptr[row + 1];
    for (int32_t k = rowptr[row] + lane; k < k1; k += TPR) acc += val[k] * x[col[k]];
  }
#pragma unroll
  for (int off = TPR / 2; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off, TPR);
  if (row < n && lane == 0) {
    double r;
    if (mode == 0) r = acc;
    else if (mode == 1) r = u[row] - acc;
    else if (mode == 2) r = u[row] + acc;
    else r = v[row] + w[row] * (u[row] - acc);
    out[row] = r;
  }
}

// out = w .* b  (the first damped Jacobi sweep from a zero guess)
__global__ void __launch_bounds__(256) mg_scale_kernel(int64_t n, const double* __restrict__ w, const double* __restrict__ b,
                                                       double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = w[i] * b[i];
}

int launch_mg_spmv(gds_ctx* ctx, int tpr, int64_t n, const int32_t* rowptr, const int32_t* col, const double* val,
                   const double* x, const double* u, const double* v, const double* w, double* out, int mode) {
  if (n <= 0) return 0;
  const int64_t blocks = (n * tpr + 255) / 256;
  if (blocks > 0x7fffffffLL) GDS_FAIL("too many rows for one launch");
  switch (tpr) {
    case 1: mg_spmv_kernel<1><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    case 2: mg_spmv_kernel<2><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    case 4: mg_spmv_kernel<4><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    case 8: mg_spmv_kernel<8><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    case 16: mg_spmv_kernel<16><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    case 32: mg_spmv_kernel<32><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, rowptr, col, val, x, u, v, w, out, mode); break;
    default: GDS_FAIL("threads per row must be a power of two up to 32");
  }
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

int launch_mg_scale(gds_ctx* ctx, int64_t n, const double* w, const double* b, double* out) {
  if (n <= 0) return 0;
  mg_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, w, b, out);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// scalars of the preconditioned iteration (one block): S[0] = r.z, S[1] = p.Ap, S[2] = alpha, S[3] = beta, S[4] = |b|^2,
// S[5] = target for |r|^2, S[6] = finished, S[7] = |r|^2
//   mode 0: p.Ap -> alpha      mode 1: |r|^2 (finished when at the target)      mode 2: r.z -> beta
//   mode 3 (start): |r|^2 = |b|^2      mode 4 (start): r.z
__global__ void __launch_bounds__(CG_THREADS) pcg_final_kernel(const double* __restrict__ partial, double* __restrict__ S,
                                                               int mode) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < CG_BLOCKS; i += CG_THREADS) acc += partial[i];
  __shared__ double total;
  block_reduce_store(acc, &total);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double v = total;
    if (mode == 0) {
      S[1] = v;
      const bool live = S[6] == 0.0 && v > 0.0 && isfinite(v);
      if (!live) S[6] = 1.0;
      S[2] = live ? S[0] / v : 0.0;
    } else if (mode == 1) {
      S[7] = v;
      if (!(v > S[5])) S[6] = 1.0;
    } else if (mode == 2) {
      S[3] = (S[6] == 0.0 && S[0] != 0.0) ? v / S[0] : 0.0;
      S[0] = v;
    } else if (mode == 3) {
      S[7] = v; S[4] = v;
    } else {
      S[0] = v;
    }
  }
}

int launch_pcg_final(gds_ctx* ctx, const double* partial, double* S, int mode) {
  pcg_final_kernel<<<1, CG_THREADS, 0, ctx->stream>>>(partial, S, mode);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

int launch_cg(gds_ctx* ctx, int what, double* x, double* r, double* p, double* ap, int64_t n, double* S, double* partial,
              int mode) {
  switch (what) {
    case 0: cg_dot_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(p, ap, n, partial); break;      // any two vectors
    case 1: cg_final_kernel<<<1, CG_THREADS, 0, ctx->stream>>>(partial, S, mode); break;
    case 2: cg_update_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(x, r, p, ap, n, S, partial); break;
    case 3: cg_direction_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(p, r, n, S); break;
    // fused iteration: `partial` holds 2 x 2048 slots (p.Ap sums, then r.r sums)
    case 4: cg_dot_commit_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(p, ap, n, partial, S); break;
    case 5: cg_update_fused_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(x, r, p, ap, n, S, partial, partial + 2048); break;
    default: cg_direction_fused_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(p, r, n, S, partial + 2048); break;
  }
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// scalar reductions of a device vector (derived observables: energies, norms, extrema -- gds/types.py:95-105 projections
// such as `(v.y ** 2).sum()`, examples/lid_driven_cavity.py:95-101): fixed grid, fixed tree => bitwise reproducible.
// kind 0 sum, 1 max, 2 min.  scratch[0 .. CG_BLOCKS) partial results, scratch[CG_BLOCKS] the result.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double red_op(int kind, double a, double b) {
  return kind == 0 ? a + b : (kind == 1 ? fmax(a, b) : fmin(a, b));
}
__device__ __forceinline__ double red_identity(int kind) {
  return kind == 0 ? 0.0 : (kind == 1 ? -INFINITY : INFINITY);
}
__device__ __forceinline__ void block_reduce_kind(double v, double* out, int kind) {
  __shared__ double sh[CG_THREADS];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = CG_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = red_op(kind, sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}
__global__ void __launch_bounds__(CG_THREADS) reduce_kernel(const double* __restrict__ x, int64_t n, int kind,
                                                            double* __restrict__ partial) {
  double acc = red_identity(kind);
  for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)CG_BLOCKS * CG_THREADS)
    acc = red_op(kind, acc, ld_stream_f64(x + i));
  block_reduce_kind(acc, partial + blockIdx.x, kind);
}
__global__ void __launch_bounds__(CG_THREADS) reduce_final_kernel(double* __restrict__ scratch, int kind) {
  double acc = red_identity(kind);
  for (int i = threadIdx.x; i < CG_BLOCKS; i += CG_THREADS) acc = red_op(kind, acc, scratch[i]);
  block_reduce_kind(acc, scratch + CG_BLOCKS, kind);
}

int launch_reduce(gds_ctx* ctx, const double* x, int64_t n, int kind, double* scratch) {
  reduce_kernel<<<CG_BLOCKS, CG_THREADS, 0, ctx->stream>>>(x, n, kind, scratch);
  GDS_CUDA(cudaGetLastError());
  reduce_final_kernel<<<1, CG_THREADS, 0, ctx->stream>>>(scratch, kind);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches += 2;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// peer-to-peer halo exchange over NVLink, inside the captured step.
// push: stores this rank's boundary values straight into the neighbours' ghost rows (peer-mapped memory), fences, and
// publishes the exchange number in each neighbour's arrival flag (deep partitions: after a ready-flag handshake).
// wait: right after the push, one thread per neighbour polls the local flag slot until the neighbour's values of the
// same exchange have arrived; only then may the next stage gather from the ghost rows.
// ---------------------------------------------------------------------------------------------------
// A wait that has timed out (a peer stopped, or the ranks enqueued different numbers of steps) raises the error flag --
// in device memory for the kernels that follow, in mapped host memory for the host, which polls it without a
// synchronisation -- and every later push / wait of this system returns at once instead of spinning again: a protocol
// failure costs ONE timeout, and the next host call on the system raises (gds_system_p2p_error).
__device__ __forceinline__ void p2p_raise(int64_t* err, int64_t* err_host, int64_t code) {
  *err = code;
  *reinterpret_cast<volatile int64_t*>(err_host) = code;
  __threadfence_system();
}

// Handshake before writing (deep partitions, whose passes also run over ghost rows): the neighbour's own stage kernel may
// still be storing (meaningless) values into its ghost rows, so every rank first announces "my stage of this exchange is
// done" (ready flag, slot 64 + rank) to its neighbours and waits for theirs; the push proper follows in stream order.
__global__ void ready_kernel(const __grid_constant__ PushParams p) {
  const int q = threadIdx.x;
  if (q >= p.n_peers) return;
  const int64_t e = *p.epoch + 1;
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p.flag[q] + 64), "l"(e) : "memory");
  if (*reinterpret_cast<volatile int64_t*>(p.err) != 0) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  int64_t seen;
  do {
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(p.ready[q]) : "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > p.timeout_ns) { p2p_raise(p.err, p.err_host, 101 + q); break; }
  } while (seen < e);
}

// grid = (blocks per neighbour, neighbours), SMALL blocks: the push runs beside the interior rows of the stage, whose grid
// keeps every SM full -- a 256-thread block fits into the room one retiring block leaves, a 1024-thread block would wait
// for the interior grid to drain.  The last block of a neighbour to finish publishes the exchange number.
#define PUSH_THREADS 256
__global__ void __launch_bounds__(PUSH_THREADS) push_kernel(const __grid_constant__ PushParams p) {
  const int q = blockIdx.y;
  const double* __restrict__ src = p.src;
  double* dst = p.dst[q];
  const int64_t stride = (int64_t)gridDim.x * PUSH_THREADS;
  for (int64_t i = p.s0[q] + (int64_t)blockIdx.x * PUSH_THREADS + threadIdx.x; i < p.s1[q]; i += stride)
    dst[p.dst_idx[i]] = src[p.src_idx[i]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(p.done + q, 1u);
    if (t == gridDim.x - 1) {
      p.done[q] = 0u;
      __threadfence_system();
      const int64_t e = *p.epoch + 1;        // number of this exchange; the counter is advanced by the wait that follows
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p.flag[q]), "l"(e) : "memory");
    }
  }
}

__global__ void wait_kernel(const __grid_constant__ WaitParams p) {
  // exchange number being completed = exchanges completed so far + 1 (the push just before used the same number)
  const int q = threadIdx.x;
  const int64_t expect = *p.epoch + 1;
  if (q < p.n_peers && *reinterpret_cast<volatile int64_t*>(p.err) == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int64_t seen;
    do {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(p.flag[q]) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > p.timeout_ns) { p2p_raise(p.err, p.err_host, 1 + q); break; }   // a peer stopped: do not hang the GPU
    } while (seen < expect);
  }
  __threadfence_system();
  __syncthreads();
  if (q == 0) *const_cast<int64_t*>(p.epoch) = expect;
}

int launch_push(gds_ctx* ctx, const PushParams& p) {
  if (p.n_peers <= 0) return 0;
  if (p.handshake) {
    ready_kernel<<<1, 32, 0, ctx->stream>>>(p);
    GDS_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  push_kernel<<<dim3((unsigned)p.blocks_per_peer, (unsigned)p.n_peers), PUSH_THREADS, 0, ctx->stream>>>(p);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

int launch_wait(gds_ctx* ctx, const WaitParams& p) {
  if (p.n_peers <= 0) return 0;
  wait_kernel<<<1, 32, 0, ctx->stream>>>(p);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// trajectory recording: copy the new accepted state into slot rec[step] of the trajectory buffer (no-op for -1)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) record_kernel(const double* __restrict__ y, double* const* __restrict__ traj,
                                                     const int64_t* __restrict__ rec, const int64_t* __restrict__ counter,
                                                     int64_t n) {
  const int64_t slot = rec[*counter];
  if (slot < 0) return;
  double* dst = *traj + slot * n;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) dst[i] = y[i];
}

int launch_record(gds_ctx* ctx, const double* y, double* const* traj, const int64_t* rec, const int64_t* counter, int64_t n) {
  int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 8);
  record_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, ctx->stream>>>(y, traj, rec, counter, n);
  GDS_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return 0;
}
