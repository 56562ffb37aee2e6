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
#define VI_PHASE(...) { __VA_ARGS__; } __syncthreads();
#else
#define VI_PHASE(...) for (int tid = 0; tid < nt; ++tid) { __VA_ARGS__; }
#endif

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
        if (l != 0.0) x = x + l * regs[((int64_t)r * n + i) * n + c];
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
VI_HD void vi_tri_reduce(const vi_tri_ws& S, int n, double* V, int tid, int nt) {
  (void)tid;
  const int ng = (nt / n) < 1 ? 1 : (nt / n);
  for (int j = 0; j + 2 < n; ++j) {
    const int lo = j + 1;
    VI_PHASE(
      if (tid >= lo + 1 && tid < n) { double x = S.X[tid * S.ld + j]; S.red1[tid] = x * x; }
    )
    VI_PHASE(
      if (tid >= lo && tid < n) {
        double xn2 = 0.0;
        for (int i = lo + 1; i < n; ++i) xn2 += S.red1[i];
        double alpha = S.X[lo * S.ld + j];
        double tau = 0.0; double beta = alpha; double scale = 0.0;
        if (xn2 != 0.0) {
          beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
        S.v[tid] = (tid == lo) ? 1.0 : S.X[tid * S.ld + j] * scale;
        if (tid == lo) { S.e[j] = beta; S.tau[j] = tau; S.d[j] = S.X[j * S.ld + j]; }
      }
    )
    const double tau = S.tau[j];
    if (tau == 0.0) {   // H_j = I (uniform across the CTA: tau lives in shared memory)
      VI_PHASE( if (tid >= lo && tid < n) V[(int64_t)j * n + tid] = (tid == lo) ? 1.0 : 0.0; )
      continue;
    }
    VI_PHASE(
      {
        int g = tid / n; int c = tid - g * n;
        if (g < ng && c >= lo) {
          double acc = 0.0;
          for (int i = lo + g; i < n; i += ng) acc += S.X[i * S.ld + c] * S.v[i];
          S.psum[g * n + c] = acc;
        }
        if (tid >= lo && tid < n) V[(int64_t)j * n + tid] = S.v[tid];
      }
    )
    VI_PHASE(
      if (tid >= lo && tid < n) {
        double p = 0.0;
        for (int g = 0; g < ng; ++g) p += S.psum[g * n + tid];
        p = tau * p;
        S.w[tid] = p;
        S.red1[tid] = p * S.v[tid];
        S.red2[tid] = S.v[tid] * S.yv[tid];
      }
    )
    VI_PHASE(
      if (tid >= lo && tid < n) {
        double dot = 0.0; double dot2 = 0.0;
        for (int i = lo; i < n; ++i) { dot += S.red1[i]; dot2 += S.red2[i]; }
        double a2 = -0.5 * tau * dot;
        S.w[tid] = S.w[tid] + a2 * S.v[tid];
        S.yv[tid] = S.yv[tid] - (tau * dot2) * S.v[tid];
      }
    )
    VI_PHASE(
      {
        int g = tid / n; int c = tid - g * n;
        if (g < ng && c >= lo) {
          double wc = S.w[c]; double vc = S.v[c];
          for (int i = lo + g; i < n; i += ng) {
            double x = S.X[i * S.ld + c];
            x = x - S.v[i] * wc;
            x = x - S.w[i] * vc;
            S.X[i * S.ld + c] = x;
          }
        }
      }
    )
  }
  VI_PHASE(
    if (tid == 0) {
      if (n >= 2) {
        S.d[n - 2] = S.X[(n - 2) * S.ld + (n - 2)];
        S.e[n - 2] = S.X[(n - 1) * S.ld + (n - 2)];
        S.tau[n - 2] = 0.0;
      }
      S.d[n - 1] = S.X[(n - 1) * S.ld + (n - 1)];
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
    double dot = 0.0;
    for (int i = j + 1; i < n; ++i) dot += vj[i] * c[(int64_t)i * stride];
    dot = dot * t;
    for (int i = j + 1; i < n; ++i) c[(int64_t)i * stride] -= dot * vj[i];
  }
}
