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
  // four terms per trip, loaded before they are added (same order of additions as the plain loop; the loads
  // of a trip overlap instead of each waiting behind the previous addition)
  double s = 0.0;
  for (int k = a + lane; k < b; k += 128) {
    const double t0 = f(k);
    const double t1 = (k + 32 < b) ? f(k + 32) : 0.0;
    const double t2 = (k + 64 < b) ? f(k + 64) : 0.0;
    const double t3 = (k + 96 < b) ? f(k + 96) : 0.0;
    s += t0;
    if (k + 32 < b) s += t1;
    if (k + 64 < b) s += t2;
    if (k + 96 < b) s += t3;
  }
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

#if defined(__CUDA_ARCH__)
typedef double2 vi_d2;
#else
struct vi_d2 { double x, y; };
#endif

// CTA-shared working set of one system.
struct vi_tri_ws {
  double* X;      // n x ld, ld even (rows 16-byte aligned: two columns per thread move as one 128-bit access)
  int ld;
  double* vw;     // 4 x (n + 2): per index i the quadruple (v[i], w[i], vnext[i], -) of the deferred / next reflector
  double* p;      // n + 2
  double* yv;     // n  right-hand side being transformed (ends as g = Q^T y)
  double* col;    // n  the pivot column, already carrying the deferred update
  double* red1;   // max(n, nt)
  double* red2;   // n
  double* psum;   // ng x npad partial mat-vec sums, npad = 2 * ceil(n/2)
  double* d;      // n   diagonal of T
  double* e;      // n   sub-diagonal of T (e[n-1] unused)
  double* tau;    // n
  double* sc;     // 8 scalars: [0] scale 2^-ex, [1] nonfinite flag
};

VI_HD int vi_tri_ld(int n) { return (n + 2) & ~1; }
VI_HD int vi_tri_npair(int n) { return (n + 1) / 2; }
VI_HD int vi_tri_groups(int n, int nt) { int ng = nt / vi_tri_npair(n); return ng < 1 ? 1 : ng; }
// doubles of CTA-shared storage needed besides X
VI_HD int vi_tri_aux_doubles(int n, int nt) {
  return 4 * (n + 2) + (n + 2) + 6 * n + (nt > n ? nt : n) + vi_tri_groups(n, nt) * 2 * vi_tri_npair(n) + 10;
}
// carve the auxiliary arrays out of one block of vi_tri_aux_doubles(n, nt) doubles (16-byte aligned)
VI_HD void vi_tri_carve(vi_tri_ws& S, double* aux, int n, int nt) {
  S.vw = aux; aux += 4 * (n + 2);
  S.p = aux; aux += n + 2;
  S.yv = aux; aux += n;
  S.col = aux; aux += n;
  S.red2 = aux; aux += n;
  S.d = aux; aux += n;
  S.e = aux; aux += n;
  S.tau = aux; aux += n;
  S.sc = aux; aux += 8;
  S.red1 = aux; aux += (nt > n ? nt : n);
  S.psum = aux + ((reinterpret_cast<uintptr_t>(aux) >> 3) & 1);     // 16-byte aligned (pair stores)
}

// X <- scl * (0.5 (G + G^T) + sum_r lam[r] Reg_r),  scl = 2^-exponent(max|X|);  yv <- y;  col <- X[:,0];
// (v, w) <- 0.  Returns (in ws.sc[1]) 1.0 if a non-finite entry was met.
// Optional rank-one downdate (leave-one-gate-out systems of the GCV objective, interpolate.py:332-349):
// with arow != NULL the system is built from G - wj a a^T and y - wj bj a, a = arow (one row of A).
VI_HD void vi_tri_load(const vi_tri_ws& S, int n, const double* G, const double* y, const double* regs,
                       const double* lam, int nreg, int tid, int nt,
                       const double* arow = nullptr, double wj = 0.0, double bj = 0.0) {
  (void)tid;
  const int ld = S.ld;
  VI_PHASE(
    double mx = 0.0; double bad = 0.0;
    // four elements per trip: the (up to 3 + nreg) global loads of each are independent, so their
    // latencies overlap instead of adding up
    for (int idx0 = tid; idx0 < n * n; idx0 += 4 * nt) {
      double x[4]; int ii[4]; int cc[4];
      VI_UNROLL4
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * nt;
        const bool ok = idx < n * n;
        const int i = ok ? idx / n : 0; const int c = ok ? idx - i * n : 0;
        ii[u] = ok ? i : -1; cc[u] = c;
        x[u] = 0.5 * (G[(int64_t)i * n + c] + G[(int64_t)c * n + i]);
      }
      for (int r = 0; r < nreg; ++r) {
        const double l = lam[r];
        if (l != 0.0) {
          VI_UNROLL4
          for (int u = 0; u < 4; ++u) x[u] = fma(l, regs[((int64_t)r * n + (ii[u] < 0 ? 0 : ii[u])) * n + cc[u]], x[u]);
        }
      }
      VI_UNROLL4
      for (int u = 0; u < 4; ++u) {
        if (ii[u] < 0) continue;
        double xv = x[u];
        if (arow) xv = xv - wj * (arow[ii[u]] * arow[cc[u]]);
        if (!(fabs(xv) <= 1.79769313486231570e308)) bad = 1.0;
        mx = fmax(mx, fabs(xv));
        S.X[ii[u] * ld + cc[u]] = xv;
      }
    }
    for (int i = tid; i < n; i += nt) {
      double t = y[i];
      if (arow) t = t - (wj * bj) * arow[i];
      if (!(fabs(t) <= 1.79769313486231570e308)) bad = 1.0;
      S.yv[i] = t;
      for (int c = n; c < ld; ++c) S.X[i * ld + c] = 0.0;      // padding columns
    }
    for (int i = tid; i < 4 * (n + 2); i += nt) S.vw[i] = 0.0;
    for (int i = tid; i < n + 2; i += nt) S.p[i] = 0.0;
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
      for (int idx = tid; idx < n * n; idx += nt) { int i = idx / n; int c = idx - i * n; S.X[i * ld + c] *= scl; }
  )
  VI_PHASE(
    for (int i = tid; i < n; i += nt) S.col[i] = S.X[i * ld];
  )
}

// Reflector k from the pivot column S.col (which already carries every earlier update): d[k], e[k],
// tau[k], vnext (S.vw[4 i + 2]) and row k of V.  Called by whole warps (warp-sum inside); the threads
// tid in [k+1, n) store their element.
VI_HD void vi_tri_reflector(const vi_tri_ws& S, int n, int k, double* V, int tid) {
  const int lo1 = k + 1;
  double tau = 0.0; double beta = 0.0; double scale = 0.0;
  const bool last = (k == n - 2);
  if (!last) {
    const double xn2 = vi_warp_sum(k + 2, n, tid & 31, [&](int i) { double x = S.col[i]; return x * x; });
    const double alpha = S.col[k + 1];
    beta = alpha;
    if (xn2 != 0.0) {
      const double r2 = alpha * alpha + xn2;
#if defined(__CUDA_ARCH__)
      const double ri = rsqrt(r2);
#else
      const double ri = 1.0 / sqrt(r2);
#endif
      const double nrm = r2 * ri;
      beta = -copysign(nrm, alpha);
      tau = 1.0 + fabs(alpha) * ri;                        // (beta - alpha) / beta
      scale = copysign(1.0, alpha) / (fabs(alpha) + nrm);  // 1 / (alpha - beta)
    }
  } else {
    beta = S.col[n - 1];
  }
  if (tid >= lo1 && tid < n) {
    double vv = 0.0;
    if (tau != 0.0) vv = (tid == lo1) ? 1.0 : S.col[tid] * scale;
    S.vw[4 * tid + 2] = vv;
    if (!last) V[(int64_t)k * n + tid] = (tau == 0.0 && tid == lo1) ? 1.0 : vv;
  }
  if (tid == 0) { S.d[k] = S.col[k]; S.e[k] = beta; S.tau[k] = tau; }
}

// Sub-phase of the first `nsub` threads only (whole warps), closed by a NAMED barrier among them: the
// other warps of the CTA are parked at the next CTA barrier and are not disturbed.
#if defined(__CUDA_ARCH__)
#define VI_SUBPHASE(nsub, ...) if (tid < (nsub)) { { __VA_ARGS__; } asm volatile("bar.sync 1, %0;" ::"r"(nsub) : "memory"); }
#else
#define VI_SUBPHASE(nsub, ...) for (int tid = 0; tid < (nsub) && tid < nt; ++tid) { __VA_ARGS__; }
#endif

// Reduction proper, FUSED form: the rank-2 update of reflector k-1 is deferred and applied in the same
// pass over the trailing matrix that forms the mat-vec for reflector k (one read + one write of X per
// Householder step instead of two reads + one write).  Iteration k = 0 .. n-2:
//   B   all threads, two columns each: x = X[i][c] - v[i] w[c] - w[i] v[c]; store; acc_c += x vnext[i]
//       -> CTA barrier
//   C   the ceil(n/32) warps that own a vector element, with named barriers between the sub-steps:
//       C1 p = tau * sum of partials, products | C2 dot products (warp-sum), w for reflector k, rhs update,
//       NEXT pivot column (column k+1 with reflector k applied), (v, w) <- (vnext, wnext) | C3 reflector k+1
//       -> CTA barrier
// i.e. 2 CTA barriers per step; the small O(n) section never stalls on the full CTA.  V (global or host):
// row k holds reflector k in columns k+1..n-1 (v[k+1] = 1 stored explicitly).  After the call S.d, S.e,
// S.tau, S.yv (= Q^T y) are final.
VI_HD void vi_tri_reduce(const vi_tri_ws& S, int n, double* V, int tid, int nt) {
  (void)tid;
  const int ld = S.ld;
  const int npair = vi_tri_npair(n), npad = 2 * npair;
  const int ng = vi_tri_groups(n, nt);
  const int nsub = (n + 31) & ~31;
  if (n >= 2) {
    VI_PHASE( if (tid < nsub) vi_tri_reflector(S, n, 0, V, tid); )
  }
  for (int k = 0; k + 1 < n; ++k) {
    const int lo1 = k + 1;
    const double tau = S.tau[k];
    // ---- B: deferred update of reflector k-1 fused with the mat-vec for reflector k -------------
    VI_PHASE(
      {
        const int g = tid / npair; const int cp = tid - g * npair;
        const int c0 = 2 * cp;
        if (g < ng && c0 + 1 >= lo1) {
          const double vc0 = S.vw[4 * c0]; const double wc0 = S.vw[4 * c0 + 1];
          const double vc1 = S.vw[4 * c0 + 4]; const double wc1 = S.vw[4 * c0 + 5];
          double a0 = 0.0; double a1 = 0.0;
          const int step = ng * ld;
          double* xp = S.X + (lo1 + g) * ld + c0;
          const double* q = S.vw + 4 * (lo1 + g);
          VI_UNROLL4
          for (int i = lo1 + g; i < n; i += ng) {
            const vi_d2 vwi = *reinterpret_cast<const vi_d2*>(q);
            const double vni = q[2];
            vi_d2 x = *reinterpret_cast<vi_d2*>(xp);
            x.x = x.x - vwi.x * wc0; x.x = x.x - vwi.y * vc0;
            x.y = x.y - vwi.x * wc1; x.y = x.y - vwi.y * vc1;
            *reinterpret_cast<vi_d2*>(xp) = x;
            a0 += x.x * vni; a1 += x.y * vni;
            xp += step; q += 4 * ng;
          }
          vi_d2 ps; ps.x = a0; ps.y = a1;
          *reinterpret_cast<vi_d2*>(S.psum + g * npad + c0) = ps;
        }
      }
    )
    // ---- C1: p = tau * (X v), products ---------------------------------------------------------
    VI_SUBPHASE(nsub,
      if (tid >= lo1 && tid < n) {
        double p = 0.0;
        for (int g = 0; g < ng; ++g) p += S.psum[g * npad + tid];
        p = tau * p;
        const double vn = S.vw[4 * tid + 2];
        S.p[tid] = p;
        S.red1[tid] = p * vn;
        S.red2[tid] = vn * S.yv[tid];
      }
    )
    // ---- C2: dot products, w, rhs, next pivot column, rotate (v, w) <- (vnext, wnext) ------------
    VI_SUBPHASE(nsub,
      {
        const double dot = vi_warp_sum(lo1, n, tid & 31, [&](int i) { return S.red1[i]; });
        const double dot2 = vi_warp_sum(lo1, n, tid & 31, [&](int i) { return S.red2[i]; });
        if (tid >= lo1 && tid < n) {
          const double a2 = -0.5 * tau * dot;
          const double vn = S.vw[4 * tid + 2];
          const double wn = S.p[tid] + a2 * vn;
          S.yv[tid] = S.yv[tid] - (tau * dot2) * vn;
          // column k+1 of the matrix with reflector k applied (vnext[k+1] = 1 when tau != 0)
          const double wlo = S.p[lo1] + a2 * S.vw[4 * lo1 + 2];
          const double vlo = S.vw[4 * lo1 + 2];
          S.col[tid] = (S.X[tid * ld + lo1] - vn * wlo) - wn * vlo;
          S.vw[4 * tid] = vn;
          S.vw[4 * tid + 1] = wn;
        }
      }
    )
    // ---- C3: reflector k + 1 (its vnext must not overwrite vw[.][2] before every C2 thread read it:
    // the named barrier above orders that) ---------------------------------------------------------
    VI_PHASE(
      if (tid < nsub && k + 2 < n) vi_tri_reflector(S, n, k + 1, V, tid);
    )
  }
  VI_PHASE(
    if (tid == 0) {
      if (n == 1) S.d[0] = S.X[0];
      else S.d[n - 1] = S.col[n - 1];
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
