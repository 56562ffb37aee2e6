// Stage 1 of the two-stage tridiagonalisation: dense symmetric X -> band of half-width 8, one CTA per system,
// everything that is O(n^3) on the FP64 tensor pipe (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4).
//
// Role on the hot path (kernel K3a): the reference solves (A^T W A + lambda R) C = A^T W b with
// scipy.linalg.lstsq = LAPACK gelsd, rcond = eps (interpolate.py:462) some 70 times per record; here every such
// system is reduced  X -> band (this file) -> tridiagonal (vi_chase.h) -> eigenvalues + rotation tape (vi_tql.h)
// and the truncated spectral solve is applied to g = Q^T y.  The one-stage Householder reduction this replaces
// (vi_tridiag_packed.h) spends 143 dependent steps of BLAS-2 work per system and sits at 6 % of the FP64 peak; the
// blocked form below needs 17 panel steps whose matrix work is two tensor-core contractions each.
//
// Storage: lower triangle of X in 8 x 8 blocks, block (I, J), I >= J, at vi_bnd_blk(I, J) * 64 doubles, element (r, c)
// at vi_bnd_el(r, c): an xor-swizzled layout in which the three fragment shapes the kernel needs (A-operand of a
// block, A-operand of its transpose, accumulator) are all shared-memory bank-conflict free (DESIGN.md has the
// derivation).  The panel factors V, W (m x 8) are column-major with leading dimension = 4 mod 8.
//
// Per panel p (block column p, rows below the diagonal block):
//   P1  warp 0: Householder QR of the m x 8 panel held in registers (one shuffle round per column)
//   P2  all warps: Y_I = sum_J A_IJ V_J for their block rows (DMMA), partial Gram blocks V^T Y, V^T V, V^T g
//   P3  all warps: T (compact WY), K = -1/2 T^T (V^T Y); W_I = (Y_I + V_I K) T (DMMA); g -= V T^T V^T g
//   P4  all warps: trailing update A_IJ -= V_I W_J^T + W_I V_J^T over the lower triangle (DMMA)
// which is  A <- Q^T A Q,  Q = I - V T V^T  (two-sided block reflector, LAPACK dsytrd/dlatrd algebra in WY form).
#pragma once
#include "vi_simt.h"

#ifndef VI_HD
#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif
#endif

#define VI_BND_NMAX 168
#ifndef VI_BND_NW
#define VI_BND_NW 6              // warps per CTA at production orders: two CTAs per SM leave 168 registers per thread,
                                 // what the register-resident panel QR needs without spilling (measured on B200 at
                                 // n = 144: 8 warps 0.72, 6 warps 0.59, 4 warps 0.60 us per system)
#endif
#define VI_BND_PART 136          // doubles per warp of partial Gram data: 64 (V^T Y) + 64 (V^T V) + 8 (V^T g)

VI_HD int vi_bnd_npad(int n) { return (n + 7) & ~7; }
VI_HD int vi_bnd_nbk(int n) { return vi_bnd_npad(n) >> 3; }
VI_HD int vi_bnd_blk(int nbk, int I, int J) { return J * nbk - (J * (J - 1)) / 2 + (I - J); }
VI_HD int vi_bnd_nblk(int n) { const int b = vi_bnd_nbk(n); return b * (b + 1) / 2; }
VI_HD int vi_bnd_el(int r, int c) { return ((c >> 2) << 5) + (((r << 2) + (c & 3)) ^ ((c >> 2) << 3)); }
VI_HD int vi_bnd_ldv(int n) { const int np = vi_bnd_npad(n); return (np > 8 ? np - 8 : 0) + 4; }
// warps of the CTA: 8 (two CTAs of 256 threads per SM at n = 144), fewer for tiny systems
VI_HD int vi_bnd_nwarp(int n) { const int b = vi_bnd_nbk(n); return b >= 9 ? VI_BND_NW : (b >= 5 ? 4 : (b >= 3 ? 2 : 1)); }
VI_HD int vi_bnd_threads(int n) { return 32 * vi_bnd_nwarp(n); }
// shared-memory doubles of one CTA
VI_HD int vi_bnd_doubles(int n, int nw = 0) {
  if (nw <= 0) nw = vi_bnd_nwarp(n);
  const int nt = 32 * nw;
  int part = nw * VI_BND_PART;
  if (part < nt) part = nt;
  return vi_bnd_nblk(n) * 64 + 2 * 8 * vi_bnd_ldv(n) + vi_bnd_npad(n) + part + 16;
}
// global storage of the block reflectors of one system: per panel p, T (64) then V column-major m x 8,
// m = npad - 8 (p + 1)
VI_HD int vi_bnd_voff(int npad, int p) { return 64 * p + 8 * p * (npad - 8) - 32 * p * (p - 1); }
VI_HD int vi_bnd_vdoubles(int n) { return vi_bnd_voff(vi_bnd_npad(n), vi_bnd_nbk(n) - 1); }
// band handed to stage 2: column j holds X[j + d][j], d = 0..8, at band[j * 9 + d]; then g (npad doubles)
VI_HD int vi_bnd_band_doubles(int n) { return 10 * vi_bnd_npad(n); }

struct vi_bnd_ws {
  double* X;      // blocks
  double* V;      // 8 x ldv, row i of the matrix at index i - 8
  double* W;      // 8 x ldv: Y, then W
  double* g;      // npad  right-hand side being transformed
  double* part;   // nwarp x VI_BND_PART (aliases the max-reduction scratch of the load phase)
  double* tau;    // 8
  double* sc;     // 8: [0] scale 2^-ex, [1] non-finite flag
  int n, npad, nbk, ldv, nw;
};

VI_HD void vi_bnd_carve(vi_bnd_ws& S, double* mem, int n, int nw = 0) {
  S.n = n; S.npad = vi_bnd_npad(n); S.nbk = S.npad >> 3; S.ldv = vi_bnd_ldv(n); S.nw = nw > 0 ? nw : vi_bnd_nwarp(n);
  const int nt = 32 * S.nw;
  int part = S.nw * VI_BND_PART;
  if (part < nt) part = nt;
  S.X = mem; mem += vi_bnd_nblk(n) * 64;
  S.V = mem; mem += 8 * S.ldv;
  S.W = mem; mem += 8 * S.ldv;
  S.g = mem; mem += S.npad;
  S.part = mem; mem += part;
  S.tau = mem; mem += 8;
  S.sc = mem; mem += 8;
}

#if defined(__CUDACC__) || defined(VI_EMU)

// X <- scl (0.5 (G + G^T) + sum_r lam[r] Reg_r [- wj a a^T]) in block layout, zero padding, scl = 2^-exponent(max|X|);
// g <- y [- wj bj a].  Same arithmetic per element as vi_trp_load (GCV downdate: interpolate.py:332-349).
// Blocks of block column J are dealt to the warps round robin; a lane covers 2 of the 64 elements of a block
// ((row lane/8 + 4 h, column lane%8): 64-byte segments of row-major G), 32-bit index arithmetic throughout.
VI_DEV void vi_bnd_load(const vi_bnd_ws& S, const double* G, const double* y, const double* regs, const double* lam,
                        int nreg, const double* arow, double wj, double bj) {
  const int n = S.n, nbk = S.nbk, tid = vi_tid(), nt = vi_nthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const double big = 1.79769313486231570e308;
  double mx = 0.0;                       // +inf marks a non-finite entry
  const int r = lane >> 3, c = lane & 7;
  const int e0 = vi_bnd_el(r, c), e1 = vi_bnd_el(r + 4, c);
  double l0 = 0.0, l1 = 0.0;             // (two regularisers cover every configuration of the reference)
  if (nreg > 0) l0 = lam[0];
  if (nreg > 1) l1 = lam[1];
  const int nn = n * n;
  for (int J = 0; J < nbk; ++J) {
    const int k = 8 * J + c;
    double* blk = S.X + vi_bnd_blk(nbk, J + warp, J) * 64;
    for (int I = J + warp; I < nbk; I += nw, blk += 64 * nw) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int h = 0; h < 2; ++h) {
        const int i = 8 * I + r + 4 * h;
        double x = 0.0;
        if (i < n && k < n) {
          const int ik = i * n + k, ki = k * n + i;
          x = 0.5 * (G[ik] + G[ki]);
          if (l0 != 0.0) x = fma(l0, regs[ik], x);
          if (l1 != 0.0) x = fma(l1, regs[nn + ik], x);
          for (int q = 2; q < nreg; ++q) {
            const double l = lam[q];
            if (l != 0.0) x = fma(l, regs[q * nn + ik], x);
          }
          if (arow) x = x - wj * (arow[i] * arow[k]);
          mx = (fabs(x) <= big) ? fmax(mx, fabs(x)) : INFINITY;
        }
        blk[h ? e1 : e0] = x;
      }
    }
  }
  for (int i = tid; i < S.npad; i += nt) {
    double t = 0.0;
    if (i < n) {
      t = y[i];
      if (arow) t = t - (wj * bj) * arow[i];
      if (!(fabs(t) <= big)) mx = INFINITY;
    }
    S.g[i] = t;
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, vi_shfl_xor(mx, o));
  if (lane == 0) S.part[warp] = mx;
  vi_cta_sync();
  double m = 0.0;
  for (int w = 0; w < nw; ++w) m = fmax(m, S.part[w]);
  const bool bad = !(m <= big);
  int ex = 0;
  double scl = 1.0;
  if (!bad && m > 0.0) { frexp(m, &ex); scl = ldexp(1.0, -ex); }
  if (tid == 0) { S.sc[0] = scl; S.sc[1] = bad ? 1.0 : 0.0; }
  if (scl != 1.0) {
    const int tot = vi_bnd_nblk(n) * 64;
    for (int idx = tid; idx < tot; idx += nt) S.X[idx] *= scl;
  }
  vi_cta_sync();
}

// ---- P1: Householder QR of panel p by ONE warp, panel in registers ---------------------------------------------
// Lane l owns the panel rows r0 + l + 32 t, t < 4 (r0 = 8 (p + 1)), all 8 columns, in registers; a panel of more than
// 128 rows keeps its rows r0 + 128 + l in shared memory (their final place in S.V), one per lane: at n = 144 that is
// the first panel only.  Column j: one shuffle round gives every lane s_c = sum_{i > pivot} a_j[i] a_c[i] for
// c = j..7 (c = j: the squared norm); then v = (1, a_j scale), beta, tau and the update a_c -= tau (v^T a_c) v are
// local except for the pivot-row elements.  The factorisation (vi_bnd_qr_compute) touches nothing but block column
// p, tau and block (p + 1, p) -- and S.V for a tall panel -- so for panels of at most 128 rows it can run beside the
// trailing update of the previous panel; the write-back of V (vi_bnd_qr_store) comes after that update is done with V.
#define VI_BND_QT 4               // register rows per lane (template default): panels of up to 128 rows without a tail
template <int QT>
struct vi_bnd_panel { double a[8][QT]; };

template <int QT>
VI_DEV void vi_bnd_qr_compute(const vi_bnd_ws& S, int p, vi_bnd_panel<QT>& P) {
  const int lane = vi_tid() & 31;
  const int r0 = 8 * (p + 1), npad = S.npad, nbk = S.nbk, ldv = S.ldv;
  double (&a)[8][QT] = P.a;
  const int i4 = r0 + 32 * QT + lane;                    // this lane's tail row (tall panels)
  const bool tail = i4 < npad;
  double* vt = S.V + (i4 - 8);                                  // column c of the tail row at vt[c * ldv]
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int t = 0; t < QT; ++t) {
    const int i = r0 + lane + 32 * t;
    const bool in = i < npad;
    const double* blk = S.X + vi_bnd_blk(nbk, in ? (i >> 3) : p + 1, p) * 64;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) a[c][t] = in ? blk[vi_bnd_el(i & 7, c)] : 0.0;
  }
  if (tail) {
    const double* blk = S.X + vi_bnd_blk(nbk, i4 >> 3, p) * 64;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) vt[c * ldv] = blk[vi_bnd_el(i4 & 7, c)];
  }
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) {
    // rows strictly below the pivot row r0 + j: t > 0, or t == 0 and lane > j
    double s[8], x4[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = j; c < 8; ++c) x4[c] = tail ? vt[c * ldv] : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = j; c < 8; ++c) {
      double acc = (lane > j) ? a[j][0] * a[c][0] : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 1; t < QT; ++t) acc = fma(a[j][t], a[c][t], acc);
      s[c] = fma(x4[j], x4[c], acc);
    }
    const double alpha = vi_shfl(a[j][0], j);
    for (int o = 16; o > 0; o >>= 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = j; c < 8; ++c) s[c] += vi_shfl_xor(s[c], o);
    }
    double beta, tau, scale;
    vi_reflector_scalars(alpha, s[j], &beta, &tau, &scale);
    // v in place of column j (rows below the pivot); pivot element becomes beta (R's diagonal)
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < QT; ++t) {
      const bool below = (t > 0) || (lane > j);
      a[j][t] = below ? a[j][t] * scale : a[j][t];
    }
    const double v4 = x4[j] * scale;
    if (tail) vt[j * ldv] = v4;
    const double vpiv = (lane == j) ? 1.0 : ((lane > j) ? a[j][0] : 0.0);   // v on this lane's t = 0 row
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = j + 1; c < 8; ++c) {
      const double apc = vi_shfl(a[c][0], j);                   // pivot-row element of column c
      const double w = tau * (apc + scale * s[c]);              // tau v^T a_c
      a[c][0] = fma(-w, vpiv, a[c][0]);
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 1; t < QT; ++t) a[c][t] = fma(-w, a[j][t], a[c][t]);
      if (tail) vt[c * ldv] = fma(-w, v4, x4[c]);
    }
    if (lane == j) a[j][0] = beta;
    if (lane == 0) S.tau[j] = tau;
  }
  if (lane < 8) {                                              // R -> block (p + 1, p)
    double* blk = S.X + vi_bnd_blk(nbk, p + 1, p) * 64;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) blk[vi_bnd_el(lane, c)] = (lane <= c) ? a[c][0] : 0.0;
  }
}

// V (unit lower trapezoidal, zeros above) -> S.V  (the tail rows of a tall panel are there already)
template <int QT>
VI_DEV void vi_bnd_qr_store(const vi_bnd_ws& S, int p, const vi_bnd_panel<QT>& P) {
  const int lane = vi_tid() & 31;
  const int r0 = 8 * (p + 1), npad = S.npad, ldv = S.ldv;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int t = 0; t < QT; ++t) {
    const int i = r0 + lane + 32 * t;
    if (i < npad) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c) {
        const int piv = r0 + c;
        S.V[c * ldv + (i - 8)] = (i > piv) ? P.a[c][t] : (i == piv ? 1.0 : 0.0);
      }
    }
  }
}

// accumulator (row lane/4, columns 2 (lane%4) + {0,1}) -> A operand of k-step t (row lane/4, column lane%4 + 4 t)
VI_DEV void vi_bnd_c_to_a(double c0, double c1, double* a0, double* a1) {
  const int lane = vi_tid() & 31;
  const int q = lane & ~3, m = lane & 3;
  // column m + 4 t lives in lane q + (m + 4 t) / 2, element (m + 4 t) % 2 = m % 2
  const int s0 = q + (m >> 1), s1 = q + (m >> 1) + 2;
  const double x0 = vi_shfl(c0, s0), y0 = vi_shfl(c1, s0);
  const double x1 = vi_shfl(c0, s1), y1 = vi_shfl(c1, s1);
  *a0 = (m & 1) ? y0 : x0;
  *a1 = (m & 1) ? y1 : x1;
}

// ---- P2: Y_I = sum_J A_IJ V_J for this warp's block rows; partial Gram blocks ---------------------------------
VI_DEV void vi_bnd_symm(const vi_bnd_ws& S, int p) {
  const int tid = vi_tid(), warp = tid >> 5, lane = tid & 31;
  const int nbk = S.nbk, ldv = S.ldv, nw = S.nw;
  const int q4 = lane >> 2, m4 = lane & 3;
  double my0 = 0.0, my1 = 0.0, ss0 = 0.0, ss1 = 0.0, ug = 0.0;
  // operand offsets that do not depend on the block: B / transposed-A pattern  (col = lane/4, row = lane%4 + 4 t)
  const int vb = q4 * ldv + m4;                          // + 8 J - 8 + 4 t
  const int ea0 = lane, ea1 = 32 + (lane ^ 8);           // A operand of a stored block, k-steps 0 and 1
  const int et0 = vi_bnd_el(m4, q4), et1 = vi_bnd_el(m4 + 4, q4);   // A operand of its transpose
  for (int I = p + 1 + warp; I < nbk; I += nw) {
    // two accumulator pairs (k-steps 0 and 1 of every block): two independent DMMA chains instead of one
    double y0 = 0.0, y1 = 0.0, z0 = 0.0, z1 = 0.0;
    // J <= I: stored blocks (I, J), one block column apart: block index grows by nbk - J - 1 per step
    {
      int b = vi_bnd_blk(nbk, I, p + 1);
      const double* vj = S.V + vb + 8 * (p + 1) - 8;
      for (int J = p + 1; J <= I; ++J) {
        const double* blk = S.X + b * 64;
        vi_mma884(y0, y1, blk[ea0], vj[0]);
        vi_mma884(z0, z1, blk[ea1], vj[4]);
        b += nbk - J - 1;
        vj += 8;
      }
    }
    // J > I: transposes of the stored blocks (J, I), contiguous in block column I
    {
      const double* blk = S.X + (vi_bnd_blk(nbk, I, I) + 1) * 64;
      const double* vj = S.V + vb + 8 * (I + 1) - 8;
      for (int J = I + 1; J < nbk; ++J) {
        vi_mma884(y0, y1, blk[et0], vj[0]);
        vi_mma884(z0, z1, blk[et1], vj[4]);
        blk += 64;
        vj += 8;
      }
    }
    y0 += z0; y1 += z1;
    // Y_I -> shared (W panel), accumulator layout: (row q4, columns 2 m4, 2 m4 + 1)
    double* wi = S.W + 8 * I - 8 + q4;
    wi[(2 * m4) * ldv] = y0;
    wi[(2 * m4 + 1) * ldv] = y1;
    vi_warp_sync();
    // partial V_I^T Y_I and V_I^T V_I: A operand = V_I^T (same addresses as V_I as a B operand)
    const double va0 = S.V[vb + 8 * I - 8], va1 = S.V[vb + 8 * I - 8 + 4];
    const double yb0 = S.W[vb + 8 * I - 8], yb1 = S.W[vb + 8 * I - 8 + 4];
    vi_mma884(my0, my1, va0, yb0);
    vi_mma884(my0, my1, va1, yb1);
    vi_mma884(ss0, ss1, va0, va0);
    vi_mma884(ss0, ss1, va1, va1);
    // partial V_I^T g_I (lanes 0..7: one column each)
    if (lane < 8) {
      const double* vc = S.V + lane * ldv + 8 * I - 8;
      const double* gi = S.g + 8 * I;
      double acc = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int r = 0; r < 8; ++r) acc = fma(vc[r], gi[r], acc);
      ug += acc;
    }
  }
  double* pw = S.part + warp * VI_BND_PART;
  pw[q4 * 8 + 2 * m4] = my0; pw[q4 * 8 + 2 * m4 + 1] = my1;
  pw[64 + q4 * 8 + 2 * m4] = ss0; pw[64 + q4 * 8 + 2 * m4 + 1] = ss1;
  if (lane < 8) pw[128 + lane] = ug;
}

// ---- P3a (warp 0): T (compact WY), K = -1/2 T^T (V^T Y), tg = T^T (V^T g) -> slot 0 of S.part -------------------
// layout of the slot afterwards: [0, 64) T row-major, [64, 128) K row-major, [128, 136) tg
VI_DEV void vi_bnd_small(const vi_bnd_ws& S) {
  const int lane = vi_tid() & 31, nw = S.nw;
  const int q4 = lane >> 2, m4 = lane & 3;
  // sums of the partial Gram blocks: lane holds elements e = lane and lane + 32 (row-major 8 x 8)
  double my[2] = {0.0, 0.0}, sv[2] = {0.0, 0.0}, ugs = 0.0;
  for (int w = 0; w < nw; ++w) {
    const double* pw = S.part + w * VI_BND_PART;
    my[0] += pw[lane]; my[1] += pw[32 + lane];
    sv[0] += pw[64 + lane]; sv[1] += pw[96 + lane];
    if (lane < 8) ugs += pw[128 + lane];
  }
  vi_warp_sync();                                               // slot 0 is read by every lane before it is overwritten
  // T (upper triangular, LAPACK dlarft forward/columnwise): lane i (mod 8) builds row i,
  //   T[i][i] = tau_i,  T[i][j] = -tau_j sum_{k=i}^{j-1} T[i][k] S[k][j]  (i < j),  S = V^T V
  const int ti = lane & 7;
  double tr[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) tr[j] = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) {
    const double tj = S.tau[j];
    double acc = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < j; ++k) {
      const int e = k * 8 + j;                                 // S[k][j] -> lane e % 32, register e / 32
      const double skj = vi_shfl(e < 32 ? sv[0] : sv[1], e & 31);
      acc = fma(tr[k], skj, acc);                              // tr[k] = 0 for k < i
    }
    tr[j] = (ti == j) ? tj : ((ti < j) ? -tj * acc : 0.0);
  }
  double* scr = S.part;
  if (lane < 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) scr[lane * 8 + j] = tr[j];
  }
  vi_warp_sync();
  // K = -1/2 T^T My: A = T^T ((T^T)[r][k] = T[k][r]: the same values as the B operand of T), B = My[k][n]
  const double tb0 = scr[m4 * 8 + q4], tb1 = scr[(m4 + 4) * 8 + q4];
  double k0 = 0.0, k1 = 0.0;
  vi_mma884(k0, k1, tb0, vi_shfl(my[0], 8 * m4 + q4));           // k = m4     -> element 8 m4 + q4 < 32
  vi_mma884(k0, k1, tb1, vi_shfl(my[1], 8 * m4 + q4));           // k = m4 + 4 -> that element - 32
  scr[64 + q4 * 8 + 2 * m4] = -0.5 * k0;
  scr[64 + q4 * 8 + 2 * m4 + 1] = -0.5 * k1;
  // tg[c] = sum_k T[k][c] ug[k]: lane k (< 8) holds row k of T and ug[k]
  const double u = (lane < 8) ? ugs : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int c = 0; c < 8; ++c) {
    const double t = vi_oct_allsum((lane < 8) ? tr[c] * u : 0.0);
    if (lane == 0) scr[128 + c] = t;
  }
}

// ---- P3b (all warps): W_I = (Y_I + V_I K) T, g_I -= V_I tg; V and T to global --------------------------------------
VI_DEV void vi_bnd_wpanel(const vi_bnd_ws& S, int p, double* Vg) {
  const int tid = vi_tid(), warp = tid >> 5, lane = tid & 31;
  const int nbk = S.nbk, ldv = S.ldv, nw = S.nw, npad = S.npad;
  const int q4 = lane >> 2, m4 = lane & 3;
  const double* scr = S.part;
  const double tb0 = scr[m4 * 8 + q4], tb1 = scr[(m4 + 4) * 8 + q4];               // B operand of T
  const double kb0 = scr[64 + m4 * 8 + q4], kb1 = scr[64 + (m4 + 4) * 8 + q4];     // B operand of K
  double tg[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int c = 0; c < 8; ++c) tg[c] = scr[128 + c];
  const int m = npad - 8 * (p + 1);
  double* vgp = Vg ? Vg + vi_bnd_voff(npad, p) : nullptr;
  if (vgp && warp == 0) {                                        // T row-major
    vgp[lane] = scr[lane];
    vgp[32 + lane] = scr[32 + lane];
  }
  for (int I = p + 1 + warp; I < nbk; I += nw) {
    const int ro = 8 * I - 8;
    // U_I = Y_I + V_I K
    double* wi = S.W + ro + q4;
    double u0 = wi[(2 * m4) * ldv], u1 = wi[(2 * m4 + 1) * ldv];
    const double* vi = S.V + ro + q4;                            // A operand of V_I: (row q4, col m4 + 4 t)
    vi_mma884(u0, u1, vi[m4 * ldv], kb0);
    vi_mma884(u0, u1, vi[(m4 + 4) * ldv], kb1);
    // W_I = U_I T
    double ua0, ua1;
    vi_bnd_c_to_a(u0, u1, &ua0, &ua1);
    double w0 = 0.0, w1 = 0.0;
    vi_mma884(w0, w1, ua0, tb0);
    vi_mma884(w0, w1, ua1, tb1);
    wi[(2 * m4) * ldv] = w0;                                     // (each lane overwrites the two values it read)
    wi[(2 * m4 + 1) * ldv] = w1;
    // g_I -= V_I tg; V_I to global
    if (lane < 8) {
      double acc = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c) acc = fma(S.V[c * ldv + ro + lane], tg[c], acc);
      S.g[8 * I + lane] -= acc;
    }
    if (vgp) {
      const int c = lane >> 3, rr = lane & 7;
      vgp[64 + c * m + (8 * I + rr - 8 * (p + 1))] = S.V[c * ldv + ro + rr];
      vgp[64 + (c + 4) * m + (8 * I + rr - 8 * (p + 1))] = S.V[(c + 4) * ldv + ro + rr];
    }
  }
}

// ---- P4: A_IJ -= V_I W_J^T + W_I V_J^T over the lower triangle of the trailing matrix ---------------------------
// blocks (I, J) of block row I for J = J0 .. J1
VI_DEV void vi_bnd_update_row(const vi_bnd_ws& S, int I, int J0, int J1) {
  const int lane = vi_tid() & 31, q4 = lane >> 2, m4 = lane & 3;
  const int nbk = S.nbk, ldv = S.ldv;
  const int ro = 8 * I - 8 + q4;
  // A operands (row q4, column m4 + 4 t), negated so that the MMA subtracts
  const double nv0 = -S.V[m4 * ldv + ro], nv1 = -S.V[(m4 + 4) * ldv + ro];
  const double nw0 = -S.W[m4 * ldv + ro], nw1 = -S.W[(m4 + 4) * ldv + ro];
  int b = vi_bnd_blk(nbk, I, J0);
  // B operands: (W_J^T)[k][n] = W[8 J + n][k]  ->  W[k * ldv + 8 J - 8 + n], k = m4 + 4 t, n = q4
  const double* wj = S.W + m4 * ldv + 8 * J0 - 8 + q4;
  const double* vj = S.V + m4 * ldv + 8 * J0 - 8 + q4;
  const int ec = vi_bnd_el(q4, 2 * m4);                              // accumulator pair (16-byte aligned)
  for (int J = J0; J <= J1; ++J) {
    double* blk = S.X + b * 64 + ec;
    double c0 = blk[0], c1 = blk[1], d0 = 0.0, d1 = 0.0;            // two independent DMMA chains per block
    vi_mma884(c0, c1, nv0, wj[0]);
    vi_mma884(d0, d1, nw0, vj[0]);
    vi_mma884(c0, c1, nv1, wj[4 * ldv]);
    vi_mma884(d0, d1, nw1, vj[4 * ldv]);
    blk[0] = c0 + d0; blk[1] = c1 + d1;
    b += nbk - J - 1;
    wj += 8; vj += 8;
  }
}

// block column p + 1 of the trailing matrix (what the next panel's QR reads), all warps
VI_DEV void vi_bnd_update_col(const vi_bnd_ws& S, int p) {
  const int warp = vi_tid() >> 5, nw = S.nw, nbk = S.nbk;
  for (int I = p + 1 + warp; I < nbk; I += nw) vi_bnd_update_row(S, I, p + 1, p + 1);
}

// the rest, block columns >= p + 2, by the warps wfirst .. nw - 1.  Rows paired short + long (row p+2+q has q+1
// blocks, row nbk-1-q has mb-q): equal work per pair.
VI_DEV void vi_bnd_update_rest(const vi_bnd_ws& S, int p, int wfirst) {
  const int warp = (vi_tid() >> 5) - wfirst, nwk = S.nw - wfirst, nbk = S.nbk;
  if (warp < 0 || nwk <= 0) return;
  const int mb = nbk - (p + 2);
  for (int q = warp; 2 * q < mb; q += nwk) {
    const int Ia = p + 2 + q, Ib = nbk - 1 - q;
    vi_bnd_update_row(S, Ia, p + 2, Ia);
    if (Ib > Ia) vi_bnd_update_row(S, Ib, p + 2, Ib);
  }
}

// Whole reduction.  After the call the band (half-width 8) sits in the diagonal and first sub-diagonal blocks,
// S.g = Q1^T y, and Vg (global, may be null) holds T and V of every panel.
// Warp 0 factors panel p + 1 WHILE the other warps finish the trailing update of panel p (look-ahead): the serial
// part of a panel step (8 dependent reflectors) is off the other warps' critical path as far as the data allow.
// pstop >= 0: only the panels [0, pstop) (the trailing matrix is then fully updated and handed to a second kernel with
// a smaller footprint, vi_bnd_store_trailing / vi_bnd_load_trailing).  QT: register rows per lane of the panel QR.
template <int QT = VI_BND_QT>
VI_DEV void vi_bnd_reduce(const vi_bnd_ws& S, double* Vg, int pstop = -1) {
  const int warp = vi_tid() >> 5;
  int npan = S.nbk - 1;
  if (pstop >= 0 && pstop < npan) npan = pstop;
  if (npan <= 0) return;
  // a panel of more than 32 QT rows keeps its tail rows in S.V while it is factored: no look-ahead for those
  auto tall = [&](int p) { return S.npad - 8 * (p + 1) > 32 * QT; };
  vi_bnd_panel<QT> P;
  if (warp == 0) { vi_bnd_qr_compute<QT>(S, 0, P); vi_bnd_qr_store<QT>(S, 0, P); }
  vi_cta_sync();
  for (int p = 0; p < npan; ++p) {
    vi_bnd_symm(S, p);
    vi_cta_sync();
    if (warp == 0) vi_bnd_small(S);
    vi_cta_sync();
    vi_bnd_wpanel(S, p, Vg);
    vi_cta_sync();
    vi_bnd_update_col(S, p);
    vi_cta_sync();
    const bool more = p + 1 < npan;
    const bool ahead = more && !tall(p + 1) && S.nw > 1;
    // look-ahead: warp 0 factors panel p + 1 while the others finish the trailing update of panel p
    if (ahead && warp == 0) vi_bnd_qr_compute<QT>(S, p + 1, P);
    else vi_bnd_update_rest(S, p, ahead ? 1 : 0);
    vi_cta_sync();
    if (more) {
      if (warp == 0) {
        if (!ahead) vi_bnd_qr_compute<QT>(S, p + 1, P);
        vi_bnd_qr_store<QT>(S, p + 1, P);
      }
      vi_cta_sync();
    }
  }
}

// ---- orders beyond the shared-memory form (n > VI_BND_NMAX): X stays in global memory --------------------------------
// The same block layout and the same phases; the blocks are read and written through L2 (1 MB per system at n = 500,
// one CTA's traffic per panel step = the trailing matrix three times), the panel factors V / W, g and the Gram
// partials stay in shared memory (75 KB at n = 500: three CTAs per SM).  The panel QR works on the panel in shared
// memory (S.V), one warp, without look-ahead.
VI_HD int vi_bnd_doubles_big(int n, int nw) {
  const int nt = 32 * nw;
  int part = nw * VI_BND_PART;
  if (part < nt) part = nt;
  return 2 * 8 * vi_bnd_ldv(n) + vi_bnd_npad(n) + part + 16;
}
VI_HD void vi_bnd_carve_big(vi_bnd_ws& S, double* mem, double* Xglobal, int n, int nw) {
  S.n = n; S.npad = vi_bnd_npad(n); S.nbk = S.npad >> 3; S.ldv = vi_bnd_ldv(n); S.nw = nw;
  const int nt = 32 * nw;
  int part = nw * VI_BND_PART;
  if (part < nt) part = nt;
  S.X = Xglobal;
  S.V = mem; mem += 8 * S.ldv;
  S.W = mem; mem += 8 * S.ldv;
  S.g = mem; mem += S.npad;
  S.part = mem; mem += part;
  S.tau = mem; mem += 8;
  S.sc = mem; mem += 8;
}

// Householder QR of panel p by warp 0, the panel in S.V (its final place); same reflector convention and the same
// results up to summation order as vi_bnd_qr_compute + vi_bnd_qr_store.  Lane l owns the rows r0 + l + 32 t.
VI_DEV void vi_bnd_qr_big(const vi_bnd_ws& S, int p) {
  const int lane = vi_tid() & 31;
  const int r0 = 8 * (p + 1), npad = S.npad, nbk = S.nbk, ldv = S.ldv;
  for (int i = r0 + lane; i < npad; i += 32) {
    const double* blk = S.X + vi_bnd_blk(nbk, i >> 3, p) * 64;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) S.V[c * ldv + (i - 8)] = blk[vi_bnd_el(i & 7, c)];
  }
  vi_warp_sync();
  for (int j = 0; j < 8; ++j) {
    const int piv = r0 + j;
    double s[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) s[c] = 0.0;
    for (int i = r0 + lane; i < npad; i += 32) {
      if (i <= piv) continue;
      const double aj = S.V[j * ldv + (i - 8)];
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c)
        if (c >= j) s[c] = fma(aj, S.V[c * ldv + (i - 8)], s[c]);
    }
    for (int o = 16; o > 0; o >>= 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c) s[c] += vi_shfl_xor(s[c], o);
    }
    const double alpha = S.V[j * ldv + (piv - 8)];
    double beta, tau, scale;
    vi_reflector_scalars(alpha, s[j], &beta, &tau, &scale);
    double w[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) w[c] = (c > j) ? tau * (S.V[c * ldv + (piv - 8)] + scale * s[c]) : 0.0;
    vi_warp_sync();                                             // pivot-row reads above, writes below
    for (int i = r0 + lane; i < npad; i += 32) {
      if (i < piv) continue;
      if (i == piv) {
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int c = 0; c < 8; ++c)
          if (c > j) S.V[c * ldv + (i - 8)] -= w[c];            // v = 1 on the pivot row
        S.V[j * ldv + (i - 8)] = beta;
      } else {
        const double v = S.V[j * ldv + (i - 8)] * scale;
        S.V[j * ldv + (i - 8)] = v;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int c = 0; c < 8; ++c)
          if (c > j) S.V[c * ldv + (i - 8)] = fma(-w[c], v, S.V[c * ldv + (i - 8)]);
      }
    }
    if (lane == 0) S.tau[j] = tau;
    vi_warp_sync();
  }
  // R -> block (p + 1, p); V unit lower trapezoidal in place
  if (lane < 8) {
    double* blk = S.X + vi_bnd_blk(nbk, p + 1, p) * 64;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) blk[vi_bnd_el(lane, c)] = (lane <= c) ? S.V[c * ldv + (r0 + lane - 8)] : 0.0;
  }
  vi_warp_sync();
  if (lane < 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c)
      if (lane <= c) S.V[c * ldv + (r0 + lane - 8)] = (lane == c) ? 1.0 : 0.0;
  }
}

VI_DEV void vi_bnd_reduce_big(const vi_bnd_ws& S, double* Vg) {
  const int warp = vi_tid() >> 5;
  const int npan = S.nbk - 1;
  if (npan <= 0) return;
  if (warp == 0) vi_bnd_qr_big(S, 0);
  vi_cta_sync();
  for (int p = 0; p < npan; ++p) {
    vi_bnd_symm(S, p);
    vi_cta_sync();
    if (warp == 0) vi_bnd_small(S);
    vi_cta_sync();
    vi_bnd_wpanel(S, p, Vg);
    vi_cta_sync();
    vi_bnd_update_col(S, p);
    vi_bnd_update_rest(S, p, 0);
    vi_cta_sync();
    if (p + 1 < npan) {
      if (warp == 0) vi_bnd_qr_big(S, p + 1);
      vi_cta_sync();
    }
  }
}

// band[j * 9 + d] = X[j + d][j] (d = 0..8, zero beyond the matrix) for the columns j < ncols, g[i] for i < ncols to
// gout.  The whole system: ncols = npad, gout = band + 9 npad.
VI_DEV void vi_bnd_store_band(const vi_bnd_ws& S, double* band, double* gout, int ncols) {
  const int n = S.n, nbk = S.nbk, tid = vi_tid(), nt = vi_nthreads();
  for (int e = tid; e < 9 * ncols; e += nt) {
    const int j = e / 9, d = e - 9 * j;
    const int i = j + d;
    double x = 0.0;
    if (i < n && j < n) x = S.X[vi_bnd_blk(nbk, i >> 3, j >> 3) * 64 + vi_bnd_el(i & 7, j & 7)];
    band[e] = x;
  }
  for (int i = tid; i < ncols; i += nt) gout[i] = S.g[i];
}
VI_DEV void vi_bnd_store_band(const vi_bnd_ws& S, double* band) {
  vi_bnd_store_band(S, band, band + 9 * S.npad, S.npad);
}

// Hand-over between the two kernels of a split reduction: after the panels [0, p1) the trailing matrix
// X[8 p1 .., 8 p1 ..] is an independent problem of order n - 8 p1 (panel p1 + q of the whole = panel q of it: the
// global reflector storage continues at vi_bnd_voff(npad, p1), the band at column 8 p1).  Blocks keep their element
// layout; Xt: vi_bnd_trailing_doubles(n, p1) doubles (blocks of the sub-problem in its own packed order, then g).
VI_HD int vi_bnd_trailing_doubles(int n, int p1) { const int m = vi_bnd_npad(n) - 8 * p1; return vi_bnd_nblk(m) * 64 + m; }
VI_DEV void vi_bnd_store_trailing(const vi_bnd_ws& S, int p1, double* Xt) {
  const int nbk = S.nbk, nb2 = nbk - p1, tid = vi_tid(), nt = vi_nthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  int c = 0;
  for (int J = p1; J < nbk; ++J)
    for (int I = J; I < nbk; ++I, ++c) {
      if (c % nw != warp) continue;
      const double* src = S.X + vi_bnd_blk(nbk, I, J) * 64;
      double* dst = Xt + vi_bnd_blk(nb2, I - p1, J - p1) * 64;
      dst[lane] = src[lane];
      dst[lane + 32] = src[lane + 32];
    }
  double* gt = Xt + vi_bnd_nblk(8 * nb2) * 64;
  for (int i = tid; i < 8 * nb2; i += nt) gt[i] = S.g[8 * p1 + i];
}
// S carved for the trailing order n - 8 p1
VI_DEV void vi_bnd_load_trailing(const vi_bnd_ws& S, const double* Xt) {
  const int tid = vi_tid(), nt = vi_nthreads();
  const int tot = vi_bnd_nblk(S.n) * 64;
  for (int i = tid; i < tot; i += nt) S.X[i] = Xt[i];
  for (int i = tid; i < S.npad; i += nt) S.g[i] = Xt[tot + i];
  if (tid == 0) { S.sc[0] = 1.0; S.sc[1] = 0.0; }
  vi_cta_sync();
}

// u <- Q1 u by one warp: block reflectors I - V T V^T in reverse panel order (u: shared memory, n entries valid,
// padded to npad with zeros by the caller or simply not read: rows >= n of V are zero).
VI_DEV void vi_bnd_apply_q(double* u, int n, const double* Vg) {
  const int lane = vi_tid() & 31;
  const int npad = vi_bnd_npad(n), nbk = npad >> 3;
  for (int p = nbk - 2; p >= 0; --p) {
    const int r0 = 8 * (p + 1), m = npad - r0;
    const double* T = Vg + vi_bnd_voff(npad, p);
    const double* V = T + 64;
    double t[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int c = 0; c < 8; ++c) t[c] = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double ui = (r0 + i < n) ? u[r0 + i] : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c) t[c] = fma(V[c * m + i], ui, t[c]);
    }
    for (int o = 16; o > 0; o >>= 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = 0; c < 8; ++c) t[c] += vi_shfl_xor(t[c], o);
    }
    double t2[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int a = 0; a < 8; ++a) {
      double acc = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int c = a; c < 8; ++c) acc = fma(T[a * 8 + c], t[c], acc);
      t2[a] = acc;
    }
    for (int i = lane; i < m; i += 32) {
      if (r0 + i < n) {
        double acc = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int c = 0; c < 8; ++c) acc = fma(V[c * m + i], t2[c], acc);
        u[r0 + i] -= acc;
      }
    }
    vi_warp_sync();
  }
}

#endif  // device / emulator
