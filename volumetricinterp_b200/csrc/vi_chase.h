// Stage 2 of the two-stage tridiagonalisation: symmetric band (half-width 8) -> tridiagonal by bulge chasing
// (Schwarz / Murata-Horikoshi / Lang's one-column-per-sweep scheme), ONE WARP per system.
//
// Sweep s annihilates column s below its sub-diagonal with a Householder reflector of length 8 acting on rows
// s+1..s+8; applying it from both sides pushes a bulge 8 rows down, whose first column the next reflector of the
// same sweep removes, and so on to the end of the matrix (step k of sweep s acts on rows s+1+8k .. s+8+8k).  The
// remaining bulge columns stay: the working band has 15 sub-diagonals.  A step only ever writes its own 8 rows
// (the update of the block below is deferred into the next step of the same sweep), so step k of sweep s+1 is
// safe once sweep s has completed its step k+1: up to four consecutive sweeps are in flight at a time, one per
// group of 8 lanes (lane = row of the 8 x 8 blocks), in lock step, two steps apart.
//
// The reflector of every step goes to global memory for the back-transformation (vi_chs_apply_q).
//
// Reference call being replaced: scipy.linalg.lstsq at interpolate.py:462 (see vi_band.h).
#pragma once
#include "vi_simt.h"

#ifndef VI_HD
#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif
#endif

#define VI_CHS_LDB 18      // rows of the working band per column: diagonal + 15 sub-diagonals, padded so that both a
                           // column walk (stride 1) and a row walk (stride LDB - 1 = 17) are bank-conflict free

VI_HD int vi_chs_nsweeps(int n) { return n > 2 ? n - 2 : 0; }
// steps of sweep s: rows s+1 .. n-1 in groups of 8
VI_HD int vi_chs_nsteps(int n, int s) { return (n - 1 - s + 7) >> 3; }
// reflectors generated before sweep s: sum_{s' < s} ceil((n - 1 - s') / 8)
VI_HD int vi_chs_off(int n, int s) {
  // sum over m = n-1-s' from n-1 down to n-s of ceil(m / 8);  F(m) = sum_{x=1}^{m} ceil(x/8) = 8 q (q+1)/2 + r (q+1), m = 8 q + r
  const int hi = n - 1, lo = n - 1 - s;     // sum_{x=lo+1}^{hi}
  const int qh = hi >> 3, rh = hi & 7, ql = lo >> 3, rl = lo & 7;
  return (4 * qh * (qh + 1) + rh * (qh + 1)) - (4 * ql * (ql + 1) + rl * (ql + 1));
}
VI_HD int vi_chs_nrefl(int n) { return vi_chs_off(n, vi_chs_nsweeps(n)); }
// doubles of global reflector storage (8 per reflector: tau, v[1..7])
VI_HD int vi_chs_rdoubles(int n) { return 8 * vi_chs_nrefl(n); }
// shared-memory doubles per system (warp): working band + g
VI_HD int vi_chs_doubles(int n) { const int np = (n + 7) & ~7; return VI_CHS_LDB * np + np + 8; }

#if defined(__CUDACC__) || defined(VI_EMU)

// Bw[c * LDB + (i - c)], i >= c.  band: 9 doubles per column + g behind (vi_bnd_store_band).
VI_DEV void vi_chs_load(double* Bw, double* g, const double* band, int n) {
  const int lane = vi_tid() & 31, np = (n + 7) & ~7;
  for (int e = lane; e < VI_CHS_LDB * np; e += 32) {
    const int c = e / VI_CHS_LDB, d = e - c * VI_CHS_LDB;
    Bw[e] = (d <= 8) ? band[c * 9 + d] : 0.0;
  }
  for (int i = lane; i < np + 8; i += 32) g[i] = (i < np) ? band[9 * np + i] : 0.0;
  vi_warp_sync();
}

// One chase step by the 8 lanes of a group (r = lane % 8 = row of the 8 x 8 blocks); `on` = false: the group idles but
// takes part in the shuffles.  Step k of sweep s works on rows r0 .. r0+7, r0 = s + 1 + 8 k, and on nothing else:
//   (k >= 1) the block left of the diagonal block, G = A[rows, r0-8 .. r0-1], first receives the PREVIOUS step's
//            reflector from the right (that update is deferred to here so that a step never writes below its own
//            rows: consecutive sweeps can then follow each other two steps apart instead of three);
//   the reflector is built from the column being cleared (column s for k = 0, the first column of G otherwise);
//   it is applied from the left to the other 7 columns of G, from both sides to the diagonal block, and to g.
// Every load of the step is issued before the reflector's scalar chain (rsqrt, reciprocal) starts.
// vprev / tauprev: the previous step's reflector of this sweep (in: step k-1's, out: this step's).
// refl: 8 doubles of this reflector in global memory (tau, v[1..7]).
VI_DEV double vi_chs_dot8(const double (&a)[8], const double (&b)[8]) {
  const double s0 = fma(a[1], b[1], a[0] * b[0]), s1 = fma(a[3], b[3], a[2] * b[2]);
  const double s2 = fma(a[5], b[5], a[4] * b[4]), s3 = fma(a[7], b[7], a[6] * b[6]);
  return (s0 + s1) + (s2 + s3);
}

VI_DEV void vi_chs_step(double* Bw, double* g, int n, int s, int k, bool on, double* refl, double (&vprev)[8],
                        double& tauprev, const int (&doff)[8]) {
  const int lane = vi_tid() & 31, r = lane & 7, base = lane & ~7;
  const int r0 = s + 1 + 8 * k;                 // first row / column of the diagonal block
  const int L = on ? ((n - r0 < 8) ? n - r0 : 8) : 0;       // rows of this step
  const int colx = (k == 0) ? s : r0 - 8;       // column being cleared
  const bool row = r < L;
  // ---- loads ---------------------------------------------------------------------------------------------------
  double G[8], d[8];
  double* grow = Bw + colx * VI_CHS_LDB + (r0 + r - colx);      // element (r0 + r, colx + q) at grow[q (LDB - 1)]
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 8; ++q) G[q] = (row && (k >= 1 || q == 0)) ? grow[q * (VI_CHS_LDB - 1)] : 0.0;
  const double* dblk = Bw + r0 * VI_CHS_LDB;                    // element (r0 + max(r,q), r0 + min(r,q)) at dblk[doff[q]]
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 8; ++q) d[q] = (row && q < L) ? dblk[doff[q]] : 0.0;
  const double gi = row ? g[r0 + r] : 0.0;
  // ---- deferred right-apply of the previous reflector of this sweep -------------------------------------------
  if (k >= 1 && tauprev != 0.0) {
    const double dot = tauprev * vi_chs_dot8(G, vprev);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) G[q] = fma(-dot, vprev[q], G[q]);
  }
  // ---- reflector from x = first column ----------------------------------------------------------------------------
  const double x = G[0];
  const double xn2 = vi_oct_allsum((r >= 1) ? x * x : 0.0);
  const double alpha = vi_shfl(x, base);
  double beta, tau, scale;
  vi_reflector_scalars(alpha, xn2, &beta, &tau, &scale);
  const double v = (r == 0) ? 1.0 : x * scale;            // zero for r >= L (x = 0 there)
  if (on) refl[r] = (r == 0) ? tau : v;
  double vv[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 8; ++q) vv[q] = vi_shfl(v, base + q);
  G[0] = (r == 0) ? beta : 0.0;
  // ---- left-apply to the other columns of G: w_c = tau sum_r v_r G[r][c] across the group's lanes --------------
  // (every lane of the warp takes part in the shuffles, whatever its group's step: a group at k = 0 carries zeros)
  {
    double w[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 1; q < 8; ++q) w[q] = v * G[q];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int o = 1; o < 8; o <<= 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int q = 1; q < 8; ++q) w[q] += vi_shfl_xor(w[q], o);
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 1; q < 8; ++q) G[q] = fma(-(tau * w[q]), v, G[q]);
  }
  if (row) {
    if (k >= 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int q = 0; q < 8; ++q) grow[q * (VI_CHS_LDB - 1)] = G[q];
    } else {
      grow[0] = G[0];
    }
  }
  // ---- two-sided on the diagonal block -------------------------------------------------------------------------------
  {
    const double pr = tau * vi_chs_dot8(d, vv);           // p = tau D v
    const double vp = vi_oct_allsum(v * pr);
    const double w = pr - (0.5 * tau * vp) * v;           // w = p - (tau/2)(v.p) v
    double ww[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) ww[q] = vi_shfl(w, base + q);
    if (tau != 0.0) {
      double* dst = Bw + r0 * VI_CHS_LDB;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int q = 0; q < 8; ++q)
        if (row && q <= r) dst[doff[q]] = (d[q] - v * ww[q]) - w * vv[q];
    }
  }
  // ---- right-hand side ---------------------------------------------------------------------------------------
  {
    const double dot = vi_oct_allsum(v * gi);
    if (row && tau != 0.0) g[r0 + r] = gi - (tau * dot) * v;
  }
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 8; ++q) vprev[q] = vv[q];
  tauprev = tau;
}

// Whole reduction by one warp.  refl: 8 * vi_chs_nrefl(n) doubles (global).  Leaves d = Bw[(j, j)], e = Bw[(j+1, j)].
VI_DEV void vi_chs_reduce(double* Bw, double* g, int n, double* refl) {
  const int lane = vi_tid() & 31, grp = lane >> 3;
  const int nsw = vi_chs_nsweeps(n);
  int s = grp;                       // this group's current sweep (s, s + 4, ...)
  int kdone = 0;                     // completed steps of it
  double vprev[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double tauprev = 0.0;
  int doff[8];                       // diagonal-block offsets of this lane's row: min(r, q) LDB + |r - q|
  {
    const int r = lane & 7;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) doff[q] = ((r < q) ? r : q) * VI_CHS_LDB + ((r < q) ? q - r : r - q);
  }
  int roff = vi_chs_off(n, s);       // reflectors generated before this group's current sweep
  for (;;) {
    const int K = (s < nsw) ? vi_chs_nsteps(n, s) : 0;
    // predecessor sweep s - 1 belongs to group (grp + 3) % 4
    const int ps = vi_shfl_i(s, ((grp + 3) & 3) * 8);
    const int pk = vi_shfl_i(kdone, ((grp + 3) & 3) * 8);
    bool can = (s < nsw);
    if (can && s > 0) {
      if (ps < s - 1) can = false;                                     // predecessor not started yet
      else if (ps == s - 1) {
        const int Kp = vi_chs_nsteps(n, s - 1);
        const int need = (kdone + 2 < Kp) ? kdone + 2 : Kp;            // it must have finished step kdone + 1
        can = pk >= need;
      }
    }
    // any group still has work?
    int alive = (s < nsw) ? 1 : 0;
    alive |= vi_shfl_i(alive, (lane + 8) & 31);
    alive |= vi_shfl_i(alive, (lane + 16) & 31);
    if (!alive) break;
    vi_chs_step(Bw, g, n, s, kdone, can, can ? refl + 8 * (roff + kdone) : refl, vprev, tauprev, doff);
    vi_warp_sync();
    if (can) {
      if (++kdone == K) {              // next sweep of this group: s + 4
        roff += K + ((s + 1 < nsw) ? vi_chs_nsteps(n, s + 1) : 0) + ((s + 2 < nsw) ? vi_chs_nsteps(n, s + 2) : 0) +
                ((s + 3 < nsw) ? vi_chs_nsteps(n, s + 3) : 0);
        s += 4; kdone = 0;
      }
    }
  }
}

// u <- Q2 u: the reflectors in reverse order.  Reflectors of one sweep act on disjoint rows, so the four lane
// groups take every fourth one of a sweep (u in shared memory or any warp-visible array).
VI_DEV void vi_chs_apply_q(double* u, int n, const double* refl) {
  const int lane = vi_tid() & 31, grp = lane >> 3, r = lane & 7;
  for (int s = vi_chs_nsweeps(n) - 1; s >= 0; --s) {
    const int K = vi_chs_nsteps(n, s);
    const double* rs = refl + 8 * vi_chs_off(n, s);
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int k = k0 + grp;
      const bool on = k < K;
      const int r0 = s + 1 + 8 * k;
      const double hv = on ? rs[8 * k + r] : 0.0;
      const double tau = vi_shfl(hv, lane & ~7);
      const double v = (r == 0) ? 1.0 : hv;
      const bool in = on && (r0 + r < n);
      const double ui = in ? u[r0 + r] : 0.0;
      const double dot = vi_oct_allsum(in ? v * ui : 0.0);
      if (in && tau != 0.0) u[r0 + r] = ui - (tau * dot) * v;
    }
    vi_warp_sync();
  }
}

#endif  // device / emulator
