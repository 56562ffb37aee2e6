// Warp-synchronous primitives used by the two-stage tridiagonalisation (vi_band.h, vi_chase.h).
//
// Two flavours of the same names:
//   * nvcc: the real thing (shuffles, mma.sync m8n8k4 f64 == SASS DMMA.8x8x4, bar.sync);
//   * host compiler + tests/cuda_emu.h: every CUDA thread is a fiber, the primitives exchange their operands
//     through per-warp mailboxes.  That build is TEST-ONLY (tests/cpu_harness.cpp) and lets the device code below
//     run unchanged in the GPU-less build container.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VI_DEV __device__ __forceinline__
VI_DEV int vi_tid() { return (int)threadIdx.x; }
VI_DEV int vi_nthreads() { return (int)blockDim.x; }
VI_DEV void vi_cta_sync() { __syncthreads(); }
VI_DEV void vi_warp_sync() { __syncwarp(); }
VI_DEV double vi_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
VI_DEV double vi_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
VI_DEV int vi_shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// 1/sqrt(x) and 1/x for NORMAL positive x (callers guarantee x > 1e-290): hardware seed (MUFU.RSQ64H / RCP64H,
// 20+ good bits) and two Newton steps, without the range checks and slow paths of the library routines -- these sit
// on the dependent chain of every Householder reflector.
VI_DEV double vi_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double xy = x * y;
  double e = fma(-xy, y, 1.0);
  y = fma(0.5 * y, e, y);
  xy = x * y;
  e = fma(-xy, y, 1.0);
  return fma(0.5 * y, e, y);
}
VI_DEV double vi_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
// D = A(8x4) B(4x8) + C on the FP64 tensor pipe.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c/d = C[lane/4][2 (lane%4) + {0,1}].
VI_DEV void vi_mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#else
#include "cuda_emu.h"
#define VI_DEV inline
VI_DEV int vi_tid() { return emu::tid(); }
VI_DEV int vi_nthreads() { return emu::cta()->nthreads; }
VI_DEV void vi_cta_sync() { emu::syncthreads(); }
VI_DEV void vi_warp_sync() { emu::syncwarp(); }
VI_DEV double vi_shfl(double v, int src) { return emu::shfl(v, src); }
VI_DEV double vi_shfl_xor(double v, int m) { return emu::shfl(v, (emu::tid() & 31) ^ m); }
VI_DEV int vi_shfl_i(int v, int src) { return emu::shfl_i(v, src); }
VI_DEV double vi_rsqrt(double x) { return 1.0 / sqrt(x); }
VI_DEV double vi_rcp(double x) { return 1.0 / x; }
VI_DEV void vi_mma884(double& d0, double& d1, double a, double b) { emu::mma884(d0, d1, a, b, d0, d1); }
#endif

// sum over the 32 lanes, every lane gets the total (fixed order: reproducible)
VI_DEV double vi_warp_allsum(double x) {
  for (int o = 16; o > 0; o >>= 1) x += vi_shfl_xor(x, o);
  return x;
}
// sum over the aligned group of 8 lanes the caller belongs to
VI_DEV double vi_oct_allsum(double x) {
  x += vi_shfl_xor(x, 1);
  x += vi_shfl_xor(x, 2);
  x += vi_shfl_xor(x, 4);
  return x;
}

// Householder reflector H = I - tau v v^T with H (alpha, x)^T = (beta, 0)^T, v = (1, scale * x), from alpha and
// xn2 = |x|^2 (LAPACK dlarfg convention).  A tail below 1e-145 of the matrix scale (the matrices are scaled to
// max|X| in [1/2, 1)) is left alone: tau = 0, beta = alpha -- 130 orders of magnitude under the eps |X| these
// reductions are accurate to, and it keeps every operand of the fast reciprocals in the normal range.
#define VI_REFL_TINY 1e-290
VI_DEV void vi_reflector_scalars(double alpha, double xn2, double* beta, double* tau, double* scale) {
  const bool live = xn2 > VI_REFL_TINY;
  const double r2 = live ? fma(alpha, alpha, xn2) : 1.0;
  const double ri = vi_rsqrt(r2);
  const double nrm = r2 * ri;
  const double sc = copysign(vi_rcp(fabs(alpha) + nrm), alpha);
  *beta = live ? -copysign(nrm, alpha) : alpha;
  *tau = live ? fma(fabs(alpha), ri, 1.0) : 0.0;
  *scale = live ? sc : 0.0;
}
