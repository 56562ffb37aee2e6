// Householder tridiagonalisation of one symmetric system by one CTA (kernel K3a).
//
// Role on the hot path: the reference solves X C = y, X = A^T W A + sum lambda_i R_i, with
// scipy.linalg.lstsq (interpolate.py:460-462) = LAPACK gelsd with rcond = eps.  For the
// symmetrised X that is the truncated eigen-expansion (vi_tql.h).  This file reduces
// X to tridiagonal form T = Q^T X Q (LAPACK dsytd2 convention, lower triangle, Q = H_0 H_1 ...),
// applies Q^T to the right-hand side on the fly and leaves the reflectors for the back-transform.
//
// The algorithm is written as PHASES separated by CTA barriers.  Every phase is a function of
// (tid, nt) that touches the CTA-shared arrays only in a race-free way, and all reductions go
// through shared scratch summed in a fixed order (so the result is independent of the thread
// count and bit-reproducible).  On the device a phase ends with __syncthreads(); in the
// test-only CPU harness (tests/cpu_harness.cpp) the same phase bodies are executed for
// tid = 0..nt-1 in turn, which lets the GPU-less build container check this exact arithmetic
// against LAPACK.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef VI_HD
#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif
#endif

#if defined(__CUDA_ARCH__)
#define VI_UNROLL4 _Pragma("unroll 4")
#else
#define VI_UNROLL4
#endif

#if defined(__CUDA_ARCH__)
#define VI_PHASE(...) { __VA_ARGS__; } __syncthreads();
#define VI_WPHASE(...) if (tid < 32) { { __VA_ARGS__; } __syncwarp(); }      // warp 0 only, warp-level barrier
#else
#define VI_PHASE(...) for (int tid = 0; tid < nt; ++tid) { __VA_ARGS__; }
#define VI_WPHASE(...) for (int tid = 0; tid < 32 && tid < nt; ++tid) { __VA_ARGS__; }
#endif

// sum_{k in [a,b)} f(k) in a fixed order: 32 strided partial sums, then an xor butterfly.  Every lane of
// the (converged) warp gets the same bits; the host version reproduces exactly that order.
template <class F>
VI_HD double vi_warp_sum(int a, int b, int lane, F f) {
#if defined(__CUDA_ARCH__)
  double s = 0.0;
  for (int k = a + lane; k < b; k += 32) s += f(k);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
#else
  (void)lane;
  double part[32], nxt[32];
  for (int l = 0; l < 32; ++l) {
    double s = 0.0;
    for (int k = a + l; k < b; k += 32) s += f(k);
    part[l] = s;
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
    for (int l = 0; l < 32; ++l) part[l] = nxt[l];
  }
  return part[0];
#endif
}

// CTA-shared working set of one system.
struct vi_tri_ws {
  double* X;      // n x ld (ld odd: conflict-free column walks), full symmetric storage
  int ld;
  double* v;      // n  current reflector
  double* w;      // n
  double* yv;     // n  right-hand side being transformed (ends as g = Q^T y)
  double* red1;   // max(n, nt)
  double* red2;   // n
  double* psum;   // ng x n partial mat-vec sums
  double* d;      // n   diagonal of T
  double* e;      // n   sub-diagonal of T (e[n-1] unused)
  double* tau;    // n
  double* sc;     // 4 scalars: [0] scale 2^-ex, [1] nonfinite flag
};

VI_HD int vi_tri_ld(int n) { return (n & 1) ? n : n + 1; }
// doubles of CTA-shared storage needed besides X
VI_HD int vi_tri_aux_doubles(int n, int nt) { int ng = nt / n; if (ng < 1) ng = 1; return 8 * n + (nt > n ? nt : n) + ng * n + 8; }

// X <- scl * (0.5 (G + G^T) + sum_r lam[r] Reg_r),  scl = 2^-exponent(max|X|);  yv <- y.
// Returns (in ws.sc[1]) 1.0 if a non-finite entry was met.
VI_HD void vi_tri_load(const vi_tri_ws& S, int n, const double* G, const double* y, const double* regs,
                       const double* lam, int nreg, int tid, int nt) {
  (void)tid;
  VI_PHASE(
    double mx = 0.0; double bad = 0.0;
    for (int idx = tid; idx < n * n; idx += nt) {
      int i = idx / n; int c = idx - i * n;
      double x = 0.5 * (G[(int64_t)i * n + c] + G[(int64_t)c * n + i]);
      for (int r = 0; r < nreg; ++r) {
        double l = lam[r];
        if (l != 0.0) x = fma(l, regs[((int64_t)r * n + i) * n + c], x);
      }
      if (!(fabs(x) <= 1.79769313486231570e308)) bad = 1.0;
      mx = fmax(mx, fabs(x));
      S.X[i * S.ld + c] = x;
    }
    for (int i = tid; i < n; i += nt) { double t = y[i]; if (!(fabs(t) <= 1.79769313486231570e308)) bad = 1.0; S.yv[i] = t; }
    S.red1[tid] = (bad != 0.0) ? -1.0 : mx;
  )
  VI_PHASE(
    if (tid == 0) {
      double mx = 0.0; double bad = 0.0;
      for (int t = 0; t < nt; ++t) { double r = S.red1[t]; if (r < 0.0) bad = 1.0; else mx = fmax(mx, r); }
      int ex = 0;
      double scl = 1.0;
      if (bad == 0.0 && mx > 0.0) { frexp(mx, &ex); scl = ldexp(1.0, -ex); }
      S.sc[0] = scl; S.sc[1] = bad;
    }
  )
  VI_PHASE(
    double scl = S.sc[0];
    if (scl != 1.0)
      for (int idx = tid; idx < n * n; idx += nt) { int i = idx / n; int c = idx - i * n; S.X[i * S.ld + c] *= scl; }
  )
}

// Reduction proper.  V (global or host): (n x n) row j holds reflector j in columns j+1..n-1
// (v[j+1] = 1 stored explicitly).  After the call S.d, S.e, S.tau, S.yv (= Q^T y) are final.
// Per step: [every warp: column norm by warp-sum -> reflector] | [all: partial mat-vec] |
// [column owners: p, products] | [their warps: two dot products by warp-sum, w, rhs update] |
// [all: rank-2 update]  -> 5 CTA barriers, no single-warp section.
VI_HD void vi_tri_reduce(const vi_tri_ws& S, int n, double* V, int tid, int nt) {
  (void)tid;
  const int ng = (nt / n) < 1 ? 1 : (nt / n);
  const int ld = S.ld;
  for (int j = 0; j + 2 < n; ++j) {
    const int lo = j + 1;
    VI_PHASE(
      const double* col = S.X + j;
      double xn2 = vi_warp_sum(lo + 1, n, tid & 31, [&](int k) { double x = col[k * ld]; return x * x; });
      double alpha = col[lo * ld];
      double tau = 0.0; double beta = alpha; double scale = 0.0;
      if (xn2 != 0.0) {
        beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      if (tid >= lo && tid < n) {
        double vv = (tid == lo) ? 1.0 : col[tid * ld] * scale;
        if (tau == 0.0) vv = (tid == lo) ? 1.0 : 0.0;
        S.v[tid] = vv;
        V[(int64_t)j * n + tid] = vv;
      }
      if (tid == lo) { S.e[j] = beta; S.tau[j] = tau; S.d[j] = S.X[j * ld + j]; }
    )
    const double tau = S.tau[j];
    if (tau == 0.0) continue;   // H_j = I (uniform across the CTA: tau lives in shared memory)
    VI_PHASE(
      {
        int g = tid / n; int c = tid - g * n;
        if (g < ng && c >= lo) {
          const int step = ng * ld;
          const double* xp = S.X + (lo + g) * ld + c;
          const double* vp = S.v + lo + g;
          const int cnt = (n - lo - g + ng - 1) / ng;
          double a0 = 0.0; double a1 = 0.0; double a2 = 0.0; double a3 = 0.0;
          int k = 0;
          for (; k + 3 < cnt; k += 4) {
            a0 += xp[0] * vp[0];
            a1 += xp[step] * vp[ng];
            a2 += xp[2 * step] * vp[2 * ng];
            a3 += xp[3 * step] * vp[3 * ng];
            xp += 4 * step; vp += 4 * ng;
          }
          for (; k < cnt; ++k) { a0 += xp[0] * vp[0]; xp += step; vp += ng; }
          S.psum[g * n + c] = (a0 + a1) + (a2 + a3);
        }
      }
    )
    VI_PHASE(
      if (tid >= lo && tid < n) {
        double p = 0.0;
        for (int g = 0; g < ng; ++g) p += S.psum[g * n + tid];
        p = tau * p;
        const double vc = S.v[tid];
        S.w[tid] = p;
        S.red1[tid] = p * vc;
        S.red2[tid] = vc * S.yv[tid];
      }
    )
    VI_PHASE(
      if (tid < ((n + 31) & ~31)) {      // the warps that own an element: each forms both dot products itself
        const double dot = vi_warp_sum(lo, n, tid & 31, [&](int k) { return S.red1[k]; });
        const double dot2 = vi_warp_sum(lo, n, tid & 31, [&](int k) { return S.red2[k]; });
        if (tid >= lo && tid < n) {
          const double vc = S.v[tid];
          S.w[tid] = S.w[tid] + (-0.5 * tau * dot) * vc;
          S.yv[tid] = S.yv[tid] - (tau * dot2) * vc;
        }
      }
    )
    VI_PHASE(
      {
        int g = tid / n; int c = tid - g * n;
        if (g < ng && c >= lo) {
          const double wc = S.w[c]; const double vc = S.v[c];
          const int step = ng * ld;
          double* xp = S.X + (lo + g) * ld + c;
          const double* vp = S.v + lo + g;
          const double* wp = S.w + lo + g;
          const int cnt = (n - lo - g + ng - 1) / ng;
          VI_UNROLL4
          for (int k = 0; k < cnt; ++k) {
            double x = xp[0];
            x = x - vp[0] * wc;
            x = x - wp[0] * vc;
            xp[0] = x;
            xp += step; vp += ng; wp += ng;
          }
        }
      }
    )
  }
  VI_PHASE(
    if (tid == 0) {
      if (n >= 2) {
        S.d[n - 2] = S.X[(n - 2) * ld + (n - 2)];
        S.e[n - 2] = S.X[(n - 1) * ld + (n - 2)];
        S.tau[n - 2] = 0.0;
      }
      S.d[n - 1] = S.X[(n - 1) * ld + (n - 1)];
      S.e[n - 1] = 0.0; S.tau[n - 1] = 0.0;
    }
  )
}

// c <- Q c = H_0 H_1 ... H_{n-3} c, sequential (one thread per system); V rows contiguous.
// `c` is a strided per-thread vector (vi_svec semantics: element i at c[i*stride]).
VI_HD void vi_tri_backtransform(int n, const double* V, const double* tau, int64_t tau_stride,
                                double* c, int64_t stride) {
  for (int j = n - 3; j >= 0; --j) {
    double t = tau[(int64_t)j * tau_stride];
    if (t == 0.0) continue;
    const double* vj = V + (int64_t)j * n;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    int i = j + 1;
    for (; i + 3 < n; i += 4) {
      double v0 = vj[i], v1 = vj[i + 1], v2 = vj[i + 2], v3 = vj[i + 3];
      d0 += v0 * c[(int64_t)i * stride];
      d1 += v1 * c[(int64_t)(i + 1) * stride];
      d2 += v2 * c[(int64_t)(i + 2) * stride];
      d3 += v3 * c[(int64_t)(i + 3) * stride];
    }
    for (; i < n; ++i) d0 += vj[i] * c[(int64_t)i * stride];
    double dot = ((d0 + d1) + (d2 + d3)) * t;
    i = j + 1;
    for (; i + 3 < n; i += 4) {
      double v0 = vj[i], v1 = vj[i + 1], v2 = vj[i + 2], v3 = vj[i + 3];
      c[(int64_t)i * stride] -= dot * v0;
      c[(int64_t)(i + 1) * stride] -= dot * v1;
      c[(int64_t)(i + 2) * stride] -= dot * v2;
      c[(int64_t)(i + 3) * stride] -= dot * v3;
    }
    for (; i < n; ++i) c[(int64_t)i * stride] -= dot * vj[i];
  }
}
