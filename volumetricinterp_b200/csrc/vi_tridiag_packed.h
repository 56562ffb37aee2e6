// Householder tridiagonalisation of one symmetric system by one CTA, PACKED form (kernel K3a).
//
// Same role and same algorithm as vi_tridiag.h (LAPACK dsytd2 convention, the rank-2 update of reflector
// k-1 deferred into the pass that forms the mat-vec for reflector k; reference call being replaced:
// scipy.linalg.lstsq at interpolate.py:462).  What changes is the storage and the work split:
//
//   * only the lower triangle is kept, column-major by OCTETS of columns: column c stores the rows
//     8*floor(c/8) .. npad-1 contiguously (npad = n rounded up to 8).  At n = 144 that is 87.5 KB instead
//     of 166 KB, so TWO systems are resident per SM and one CTA's latency-bound vector section overlaps the
//     other's matrix pass;
//   * every stored element is touched once per Householder step by one thread, which applies the deferred
//     update and feeds BOTH the column sum (X v)_c and the row sum (X v)_i it belongs to (half the
//     shared-memory traffic and half the FMAs of the full-square form);
//   * a thread works on tiles of 2 rows x 8 columns (the 2 rows are one 128-bit access per column), a warp
//     owns one or two octets and walks their row pairs 32 at a time; the 8 column sums of an octet are
//     reduced across the warp once per step (recursive halving, fixed order), the 2 row sums of a tile go
//     to a per-tile slot that the vector section adds up.
//
// As in vi_tridiag.h the algorithm is a sequence of phases separated by barriers, and the test-only CPU
// harness (tests/cpu_harness.cpp) executes the same phase bodies thread by thread, warp reductions
// included in the same order.
#pragma once
#include "vi_tridiag.h"
#include <vector>      // host restatement of the vector section only

struct vi_trp_ws {
  double* X;      // packed lower triangle, vi_trp_xdoubles(n) doubles
  double* Pr;     // 2 x ntiles: row sums of every tile (rows 2 ip, 2 ip + 1)
  double* v;      // npad  deferred reflector k-1
  double* w;      // npad  its companion vector
  double* vn;     // npad  reflector k
  double* pcol;   // npad  column sums
  double* p;      // n
  double* yv;     // n  right-hand side being transformed (ends as g = Q^T y)
  double* col;    // n  pivot column, carrying every earlier update
  double* red1;   // max(n, nt)
  double* red2;   // n
  double* d;      // n
  double* e;      // n
  double* tau;    // n
  double* sc;     // 8 scalars: [0] scale 2^-ex, [1] non-finite flag
  double* part;   // 2 per vector-section warp: partial sums of p.vn and vn.y
  double* part2;  // 1 per vector-section warp: partial sums of |pivot column|^2
  int n, npad, noct, nwarp;
};

VI_HD int vi_trp_npad(int n) { return (n + 7) & ~7; }
VI_HD int vi_trp_noct(int n) { return vi_trp_npad(n) >> 3; }
// warps of the CTA: a little over half the octets, see vi_trp_octet_of
VI_HD int vi_trp_nwarp(int n) { const int o = vi_trp_noct(n); const int w = (o + 2) / 2; return w > o ? o : w; }
// doubles stored before octet q: sum_{t<q} 8 (npad - 8 t)
VI_HD int vi_trp_octoff(int npad, int q) { return 8 * q * npad - 32 * q * (q - 1); }
VI_HD int vi_trp_xdoubles(int n) { const int np = vi_trp_npad(n); return vi_trp_octoff(np, np >> 3); }
// element (i, c) with i >= 8 floor(c/8)
VI_HD int vi_trp_idx(int npad, int i, int c) {
  const int q = c >> 3;
  return vi_trp_octoff(npad, q) + (c - 8 * q) * (npad - 8 * q) + (i - 8 * q);
}
// tiles (row pair ip >= 4 q, octet q): slot = tileoff(q) + ip - 4 q
VI_HD int vi_trp_tileoff(int npad, int q) { return q * (npad >> 1) - 2 * q * (q - 1); }
VI_HD int vi_trp_ntiles(int n) { const int np = vi_trp_npad(n); return vi_trp_tileoff(np, np >> 3); }
VI_HD int vi_trp_threads(int n) { return 32 * vi_trp_nwarp(n); }
// doubles of CTA-shared storage, X included
VI_HD int vi_trp_doubles(int n) {
  const int np = vi_trp_npad(n), nt = vi_trp_threads(n);
  return vi_trp_xdoubles(n) + 2 * vi_trp_ntiles(n) + 4 * np + 7 * n + (nt > n ? nt : n) + 8 + 8 + 4 * ((n + 31) / 32);
}

VI_HD void vi_trp_carve(vi_trp_ws& W, double* mem, int n) {
  const int np = vi_trp_npad(n), nt = vi_trp_threads(n);
  W.n = n; W.npad = np; W.noct = np >> 3; W.nwarp = vi_trp_nwarp(n);
  W.X = mem; mem += vi_trp_xdoubles(n);          // even number of doubles: everything below stays 16-byte aligned
  W.Pr = mem; mem += 2 * vi_trp_ntiles(n);
  W.v = mem; mem += np;
  W.w = mem; mem += np;
  W.vn = mem; mem += np;
  W.pcol = mem; mem += np;
  W.p = mem; mem += n;
  W.yv = mem; mem += n;
  W.col = mem; mem += n;
  W.red2 = mem; mem += n;
  W.d = mem; mem += n;
  W.e = mem; mem += n;
  W.tau = mem; mem += n;
  W.sc = mem; mem += 8;
  W.red1 = mem; mem += (nt > n ? nt : n);
  W.part = mem; mem += 2 * ((n + 31) / 32);
  W.part2 = mem; mem += 2 * ((n + 31) / 32);
}

// X <- scl * (0.5 (G + G^T) + sum_r lam[r] Reg_r) (lower triangle, packed), scl = 2^-exponent(max|X|);
// yv <- y; col <- X[:,0]; v, w, vn <- 0.  Same arithmetic per element as vi_tri_load, including the optional
// rank-one downdate of the GCV objective (interpolate.py:332-349).
VI_HD void vi_trp_load(const vi_trp_ws& W, const double* G, const double* y, const double* regs, const double* lam,
                       int nreg, int tid, int nt, const double* arow = nullptr, double wj = 0.0, double bj = 0.0) {
  (void)tid;
  const int n = W.n, npad = W.npad;
  VI_PHASE(
    double mx = 0.0; double bad = 0.0;
    const int warp = tid >> 5; const int lane = tid & 31; const int nwarp = nt >> 5;
    // a warp takes 4 columns of one octet per trip, lanes walk the rows: shared-memory stores are
    // conflict-free (consecutive rows), G[c][i] is read coalesced, G[i][c] strided (both through L2)
    for (int cb = 4 * warp; cb < npad; cb += 4 * nwarp) {
      const int r0 = cb & ~7;
      for (int i = r0 + lane; i < npad; i += 32) {
        double x[4];
        VI_UNROLL4
        for (int u = 0; u < 4; ++u) {
          const int c = cb + u;
          const bool in = (i < n) && (c < n);
          x[u] = in ? 0.5 * (G[(int64_t)i * n + c] + G[(int64_t)c * n + i]) : 0.0;
        }
        for (int r = 0; r < nreg; ++r) {
          const double l = lam[r];
          if (l != 0.0) {
            VI_UNROLL4
            for (int u = 0; u < 4; ++u) {
              const int c = cb + u;
              if ((i < n) && (c < n)) x[u] = fma(l, regs[((int64_t)r * n + i) * n + c], x[u]);
            }
          }
        }
        VI_UNROLL4
        for (int u = 0; u < 4; ++u) {
          const int c = cb + u;
          double xv = x[u];
          if (arow && (i < n) && (c < n)) xv = xv - wj * (arow[i] * arow[c]);
          if (!(fabs(xv) <= 1.79769313486231570e308)) bad = 1.0;
          mx = fmax(mx, fabs(xv));
          W.X[vi_trp_idx(npad, i, c)] = xv;
        }
      }
    }
    for (int i = tid; i < n; i += nt) {
      double t = y[i];
      if (arow) t = t - (wj * bj) * arow[i];
      if (!(fabs(t) <= 1.79769313486231570e308)) bad = 1.0;
      W.yv[i] = t;
      W.p[i] = 0.0;
    }
    for (int i = tid; i < npad; i += nt) { W.v[i] = 0.0; W.w[i] = 0.0; W.vn[i] = 0.0; W.pcol[i] = 0.0; }
    W.red1[tid] = (bad != 0.0) ? -1.0 : mx;
  )
  VI_PHASE(
    if (tid == 0) {
      double mx = 0.0; double bad = 0.0;
      for (int t = 0; t < nt; ++t) { double r = W.red1[t]; if (r < 0.0) bad = 1.0; else mx = fmax(mx, r); }
      int ex = 0;
      double scl = 1.0;
      if (bad == 0.0 && mx > 0.0) { frexp(mx, &ex); scl = ldexp(1.0, -ex); }
      W.sc[0] = scl; W.sc[1] = bad;
    }
  )
  VI_PHASE(
    const double scl = W.sc[0];
    const int tot = vi_trp_xdoubles(n);
    if (scl != 1.0)
      for (int idx = tid; idx < tot; idx += nt) W.X[idx] *= scl;
  )
  VI_PHASE(
    for (int i = tid; i < n; i += nt) W.col[i] = W.X[i];       // column 0 starts the packed array
  )
}

// Reflector k from the pivot column W.col: d[k], e[k], tau[k], vn and row k of V (vi_tri_reflector with the
// vectors in separate arrays).  vn[k] is cleared: the tile pass reads vn for the odd row k when k + 1 is odd.
VI_HD void vi_trp_reflector(const vi_trp_ws& W, int k, double* V, int tid) {
  const int n = W.n, lo1 = k + 1;
  double tau = 0.0; double beta = 0.0; double scale = 0.0;
  const bool last = (k == n - 2);
  if (!last) {
    const double xn2 = vi_warp_sum(k + 2, n, tid & 31, [&](int i) { double x = W.col[i]; return x * x; });
    const double alpha = W.col[k + 1];
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
      tau = 1.0 + fabs(alpha) * ri;
      scale = copysign(1.0, alpha) / (fabs(alpha) + nrm);
    }
  } else {
    beta = W.col[n - 1];
  }
  if (tid >= lo1 && tid < n) {
    double vv = 0.0;
    if (tau != 0.0) vv = (tid == lo1) ? 1.0 : W.col[tid] * scale;
    W.vn[tid] = vv;
    if (!last) V[(int64_t)k * n + tid] = (tau == 0.0 && tid == lo1) ? 1.0 : vv;
  }
  if (tid == k) W.vn[k] = 0.0;
  if (tid == 0) { W.d[k] = W.col[k]; W.e[k] = beta; W.tau[k] = tau; }
}

// One tile BELOW the diagonal block of its octet: rows (2 ip, 2 ip + 1), ip >= 4 q + 4, x the 8 columns of octet
// q.  xb points at element (row 0, column 8 q) of a virtual column-major block of leading dimension
// len = npad - 8 q, so element (i, 8 q + j) is xb[j len + i].  Applies x <- x - v_i w_c - w_i v_c, accumulates
// acc[j] += x vn_i (column sums) and returns the two row sums sum_c x vn_c.
VI_HD void vi_trp_tile(double* xb, int len, const vi_trp_ws& W, int q, int ip, double* acc, double* pr) {
  const int i0 = 2 * ip;
  const vi_d2 v01 = *reinterpret_cast<const vi_d2*>(W.v + i0);
  const vi_d2 w01 = *reinterpret_cast<const vi_d2*>(W.w + i0);
  const vi_d2 n01 = *reinterpret_cast<const vi_d2*>(W.vn + i0);
  double pr0 = 0.0, pr1 = 0.0;
  const vi_d2* vq = reinterpret_cast<const vi_d2*>(W.v + 8 * q);      // column operands, two columns per load
  const vi_d2* wq = reinterpret_cast<const vi_d2*>(W.w + 8 * q);
  const vi_d2* nq = reinterpret_cast<const vi_d2*>(W.vn + 8 * q);
  // (a finished column c < lo1 of the first active octet is processed like the others: its stored values are
  // never read again, its column sum is not used and vn[c] = 0 keeps it out of the row sums)
  // One column after the other, the next column's load issued before the current one's store (the compiler
  // cannot hoist it on its own: it does not know that the columns do not overlap).  Issuing four or eight
  // loads together was measured slower on B200 (1434 vs 1395 ms per 10k records).
  double* xp = xb + i0;
  vi_d2 xn = *reinterpret_cast<vi_d2*>(xp);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j2 = 0; j2 < 4; ++j2) {
    const vi_d2 vc = vq[j2], wc = wq[j2], nc = nq[j2];
    {
      vi_d2 x = xn;
      xn = *reinterpret_cast<vi_d2*>(xp + len);
      x.x = x.x - v01.x * wc.x; x.x = x.x - w01.x * vc.x;
      x.y = x.y - v01.y * wc.x; x.y = x.y - w01.y * vc.x;
      *reinterpret_cast<vi_d2*>(xp) = x;
      acc[2 * j2] += x.x * n01.x; acc[2 * j2] += x.y * n01.y;
      pr0 += x.x * nc.x; pr1 += x.y * nc.x;
      xp += len;
    }
    {
      vi_d2 x = xn;
      if (j2 < 3) xn = *reinterpret_cast<vi_d2*>(xp + len);
      x.x = x.x - v01.x * wc.y; x.x = x.x - w01.x * vc.y;
      x.y = x.y - v01.y * wc.y; x.y = x.y - w01.y * vc.y;
      *reinterpret_cast<vi_d2*>(xp) = x;
      acc[2 * j2 + 1] += x.x * n01.x; acc[2 * j2 + 1] += x.y * n01.y;
      pr0 += x.x * nc.y; pr1 += x.y * nc.y;
      xp += len;
    }
  }
  pr[0] = pr0; pr[1] = pr1;
}

// The 8 x 8 diagonal block of octet q, one lane per (row pair rp = lane % 4, column j = lane / 4): an element
// counts once for its column if i >= c and once for its row if i > c; stored positions above the diagonal are
// dead.  Returns this lane's contribution to column sum j (cj) and to the two row sums of its row pair.
VI_HD void vi_trp_diag_lane(double* xb, int len, const vi_trp_ws& W, int q, int lo1, int lane,
                            double* cj, double* r0, double* r1) {
  const int rp = lane & 3, j = lane >> 2;      // row pair fastest: a quarter-warp reads 2 columns x 4 consecutive row pairs
  const int ip = 4 * q + rp, i0 = 2 * ip, c = 8 * q + j;
  *cj = 0.0; *r0 = 0.0; *r1 = 0.0;
  if (ip < (lo1 >> 1) || c < lo1 || c > i0 + 1) return;
  const vi_d2 v01 = *reinterpret_cast<const vi_d2*>(W.v + i0);
  const vi_d2 w01 = *reinterpret_cast<const vi_d2*>(W.w + i0);
  const vi_d2 n01 = *reinterpret_cast<const vi_d2*>(W.vn + i0);
  const double vc = W.v[c], wc = W.w[c], nc = W.vn[c];
  vi_d2* xp = reinterpret_cast<vi_d2*>(xb + j * len + i0);
  vi_d2 x = *xp;
  x.x = x.x - v01.x * wc; x.x = x.x - w01.x * vc;
  x.y = x.y - v01.y * wc; x.y = x.y - w01.y * vc;
  *xp = x;
  double s = x.y * n01.y;                        // i0 + 1 >= c here
  if (i0 >= c) s += x.x * n01.x;
  *cj = s;
  if (i0 > c) *r0 = x.x * nc;
  if (i0 + 1 > c) *r1 = x.y * nc;
}

// The octets of warp `warp` at a step whose first active octet is a: slot 0 is octet a + warp, slot 1 the octet
// folded back from the end, a + 2 nwarp - 1 - warp (long octets come first, so the warps that got the longest
// ones get the shortest of the rest, or none).  Every active octet belongs to exactly one warp.
VI_HD int vi_trp_octet_of(const vi_trp_ws& W, int warp, int slot, int a) {
  const int q = (slot == 0) ? a + warp : a + 2 * W.nwarp - 1 - warp;
  return q < W.noct ? q : -1;
}

// lane -> which of the 8 column sums it ends up holding after vi_trp_reduce8 (lanes with lane % 4 == 0 store)
VI_HD int vi_trp_reduce_slot(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

#if defined(__CUDA_ARCH__)
// sum over the 32 lanes of each of 8 values by recursive halving: 4 + 2 + 1 exchanges, then two butterfly
// steps on the single value left (fixed order -> reproducible)
__device__ __forceinline__ double vi_trp_reduce8(const double* acc, int lane) {
  double a4[4], a2[2], a1;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double send = hi ? acc[t] : acc[t + 4];
      const double keep = hi ? acc[t + 4] : acc[t];
      a4[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const double send = hi ? a4[t] : a4[t + 2];
      const double keep = hi ? a4[t + 2] : a4[t];
      a2[t] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const double send = hi ? a2[0] : a2[1];
    const double keep = hi ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a1 = a1 + __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 = a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}
#else
// host restatement: acc[lane][8] -> out[lane], same exchanges in the same order
inline void vi_trp_reduce8_host(const double (*acc)[8], double* out) {
  double a4[32][4], a2[32][2], a1[32], t1[32];
  for (int l = 0; l < 32; ++l) {
    const bool hi = l & 16;
    for (int t = 0; t < 4; ++t) {
      const double keep = hi ? acc[l][t + 4] : acc[l][t];
      const double recv = (l ^ 16) & 16 ? acc[l ^ 16][t] : acc[l ^ 16][t + 4];
      a4[l][t] = keep + recv;
    }
  }
  for (int l = 0; l < 32; ++l) {
    const bool hi = l & 8;
    for (int t = 0; t < 2; ++t) {
      const double keep = hi ? a4[l][t + 2] : a4[l][t];
      const double recv = (l ^ 8) & 8 ? a4[l ^ 8][t] : a4[l ^ 8][t + 2];
      a2[l][t] = keep + recv;
    }
  }
  for (int l = 0; l < 32; ++l) {
    const bool hi = l & 4;
    const double keep = hi ? a2[l][1] : a2[l][0];
    const double recv = (l ^ 4) & 4 ? a2[l ^ 4][0] : a2[l ^ 4][1];
    a1[l] = keep + recv;
  }
  for (int l = 0; l < 32; ++l) t1[l] = a1[l] + a1[l ^ 2];
  for (int l = 0; l < 32; ++l) out[l] = t1[l] + t1[l ^ 1];
}
#endif

// Matrix pass of step k (lo1 = k + 1) for one warp: its octets; per octet the diagonal block (one lane per row
// pair and column), then the tiles below it, 32 row pairs per trip.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void vi_trp_pass(const vi_trp_ws& W, int lo1, int warp, int lane) {
  const int npad = W.npad, nrp = npad >> 1, a = lo1 >> 3;
#pragma unroll 1
  for (int slot = 0; slot < 2; ++slot) {
    const int q = vi_trp_octet_of(W, warp, slot, a);
    if (q < 0) continue;
    const int len = npad - 8 * q;
    double* xb = W.X + vi_trp_octoff(npad, q) - 8 * q;
    double* prq = W.Pr + 2 * (vi_trp_tileoff(npad, q) - 4 * q);
    double acc[8];
    {
      double cj, r0, r1;
      vi_trp_diag_lane(xb, len, W, q, lo1, lane, &cj, &r0, &r1);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ((lane >> 2) == j) ? cj : 0.0;
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
      }
      const int ipd = 4 * q + (lane & 3);
      if (lane < 4 && ipd >= (lo1 >> 1)) {
        vi_d2 o2; o2.x = r0; o2.y = r1;
        *reinterpret_cast<vi_d2*>(prq + 2 * ipd) = o2;
      }
    }
    const int ip0 = (4 * q + 4 > (lo1 >> 1)) ? 4 * q + 4 : (lo1 >> 1);
#pragma unroll 1
    for (int ip = ip0 + lane; ip < nrp; ip += 32) {
      double pr[2];
      vi_trp_tile(xb, len, W, q, ip, acc, pr);
      vi_d2 o; o.x = pr[0]; o.y = pr[1];
      *reinterpret_cast<vi_d2*>(prq + 2 * ip) = o;
    }
    const double tot = vi_trp_reduce8(acc, lane);
    if ((lane & 3) == 0) W.pcol[8 * q + vi_trp_reduce_slot(lane)] = tot;
  }
}
#else
inline void vi_trp_pass_host(const vi_trp_ws& W, int lo1, int warp) {
  const int npad = W.npad, nrp = npad >> 1, a = lo1 >> 3;
  for (int slot = 0; slot < 2; ++slot) {
    const int q = vi_trp_octet_of(W, warp, slot, a);
    if (q < 0) continue;
    const int len = npad - 8 * q;
    double* xb = W.X + vi_trp_octoff(npad, q) - 8 * q;
    double* prq = W.Pr + 2 * (vi_trp_tileoff(npad, q) - 4 * q);
    double acc[32][8], tot[32], r0[32], r1[32], t0[32], t1[32];
    for (int lane = 0; lane < 32; ++lane) {
      double cj;
      vi_trp_diag_lane(xb, len, W, q, lo1, lane, &cj, &r0[lane], &r1[lane]);
      for (int j = 0; j < 8; ++j) acc[lane][j] = ((lane >> 2) == j) ? cj : 0.0;
    }
    for (int o = 4; o < 32; o <<= 1) {
      for (int l = 0; l < 32; ++l) { t0[l] = r0[l] + r0[l ^ o]; t1[l] = r1[l] + r1[l ^ o]; }
      for (int l = 0; l < 32; ++l) { r0[l] = t0[l]; r1[l] = t1[l]; }
    }
    for (int lane = 0; lane < 4; ++lane) {
      const int ipd = 4 * q + (lane & 3);
      if (ipd >= (lo1 >> 1)) { prq[2 * ipd] = r0[lane]; prq[2 * ipd + 1] = r1[lane]; }
    }
    const int ip0 = (4 * q + 4 > (lo1 >> 1)) ? 4 * q + 4 : (lo1 >> 1);
    for (int lane = 0; lane < 32; ++lane) {
      for (int ip = ip0 + lane; ip < nrp; ip += 32) {
        double pr[2];
        vi_trp_tile(xb, len, W, q, ip, acc[lane], pr);
        prq[2 * ip] = pr[0]; prq[2 * ip + 1] = pr[1];
      }
    }
    vi_trp_reduce8_host(acc, tot);
    for (int lane = 0; lane < 32; lane += 4) W.pcol[8 * q + vi_trp_reduce_slot(lane)] = tot[lane];
  }
}
#endif

// ---- vector section of one Householder step (threads tid < nsub, i.e. the warps that own a vector element) -----
// Every thread keeps its element of p, vn, y and of the next pivot column in registers; the three dot products
// of a step (p.vn, vn.y, |column|^2) are block sums: a butterfly inside each warp, one partial per warp in
// shared memory, a named barrier, and every thread adds the partials in warp order.  The pieces below are the
// per-thread arithmetic, shared by the device code and its host restatement.

// C1: p_i = tau (column sum + row sums of the tiles left of the diagonal); returns p_i, sets r1 = p_i vn_i, r2 = vn_i y_i
VI_HD double vi_trp_c1(const vi_trp_ws& W, int lo1, double tau, int tid, double vn_i, double yv_i, double* r1, double* r2) {
  const int npad = W.npad;
  const int qlo = lo1 >> 3, qhi = tid >> 3;
  const double* prt = W.Pr + 2 * (tid >> 1) + (tid & 1);
  // four running sums (fixed assignment q % 4 relative to qlo), combined at the end: a short dependent chain
  double s0 = W.pcol[tid], s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int q = qlo; q <= qhi; q += 4) {
    const double t0 = prt[2 * (vi_trp_tileoff(npad, q) - 4 * q)];
    const double t1 = (q + 1 <= qhi) ? prt[2 * (vi_trp_tileoff(npad, q + 1) - 4 * (q + 1))] : 0.0;
    const double t2 = (q + 2 <= qhi) ? prt[2 * (vi_trp_tileoff(npad, q + 2) - 4 * (q + 2))] : 0.0;
    const double t3 = (q + 3 <= qhi) ? prt[2 * (vi_trp_tileoff(npad, q + 3) - 4 * (q + 3))] : 0.0;
    s0 += t0; s1 += t1; s2 += t2; s3 += t3;
  }
  const double p = tau * ((s0 + s1) + (s2 + s3));
  *r1 = p * vn_i;
  *r2 = vn_i * yv_i;
  return p;
}

// C2: w_i, rhs update, element i of the next pivot column (column k+1 with reflector k applied); returns col_i
VI_HD double vi_trp_c2(const vi_trp_ws& W, int lo1, double tau, double dot, double dot2, int tid, double p_i,
                       double vn_i, double yv_i, double xcol) {
  const double a2 = -0.5 * tau * dot;
  const double wn = p_i + a2 * vn_i;
  W.yv[tid] = yv_i - (tau * dot2) * vn_i;
  const double vlo = W.vn[lo1];
  const double wlo = W.p[lo1] + a2 * vlo;
  const double col = (xcol - vn_i * wlo) - wn * vlo;
  W.col[tid] = col;
  W.v[tid] = vn_i;
  W.w[tid] = wn;
  return col;
}

// C3: reflector kk from the pivot column (col_i in a register, |col[kk+2:]|^2 = xn2 given); same arithmetic as
// vi_trp_reflector
VI_HD void vi_trp_c3(const vi_trp_ws& W, int kk, double xn2, double* V, int tid, double col_i) {
  const int n = W.n, lo = kk + 1;
  double tau = 0.0; double beta = 0.0; double scale = 0.0;
  const bool last = (kk == n - 2);
  if (!last) {
    const double alpha = W.col[kk + 1];
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
      tau = 1.0 + fabs(alpha) * ri;
      scale = copysign(1.0, alpha) / (fabs(alpha) + nrm);
    }
  } else {
    beta = W.col[n - 1];
  }
  if (tid >= lo && tid < n) {
    double vv = 0.0;
    if (tau != 0.0) vv = (tid == lo) ? 1.0 : col_i * scale;
    W.vn[tid] = vv;
    if (!last) V[(int64_t)kk * n + tid] = (tau == 0.0 && tid == lo) ? 1.0 : vv;
  }
  if (tid == kk) W.vn[kk] = 0.0;
  if (tid == 0) { W.d[kk] = W.col[kk]; W.e[kk] = beta; W.tau[kk] = tau; }
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double vi_bfly(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
#define VI_NAMED_BAR(nsub) asm volatile("bar.sync 1, %0;" ::"r"(nsub) : "memory")
#else
inline void vi_bfly_host(const double* x, double* out) {      // 32 lanes -> the value every lane ends with
  double a[32], b[32];
  for (int l = 0; l < 32; ++l) a[l] = x[l];
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) b[l] = a[l] + a[l ^ o];
    for (int l = 0; l < 32; ++l) a[l] = b[l];
  }
  *out = a[0];
}
#endif

// Reduction proper.  Per Householder step: matrix pass (all warps) -> CTA barrier -> vector section on the
// ceil(n/32) warps that own a vector element (named barriers between its three sub-steps) -> CTA barrier.
// After the call W.d, W.e, W.tau, W.yv (= Q^T y) are final; V row k holds reflector k in columns k+1..n-1.
VI_HD void vi_trp_reduce(const vi_trp_ws& W, double* V, int tid, int nt) {
  (void)tid;
  const int n = W.n, npad = W.npad;
  const int nsub = (n + 31) & ~31;
  if (n >= 2) {
    VI_PHASE( if (tid < nsub) vi_trp_reflector(W, 0, V, tid); )
  }
#if defined(__CUDA_ARCH__) && defined(VI_TRP_PROFILE)
  long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define VI_TRP_TICK(i) { const long long now_ = clock64(); pt[i] += now_ - t_; t_ = now_; }
  long long t_ = clock64();
#else
#define VI_TRP_TICK(i)
#endif
  for (int k = 0; k + 1 < n; ++k) {
    const int lo1 = k + 1;
    const double tau = W.tau[k];
    // ---- B: deferred update of reflector k-1 fused with both halves of the symmetric mat-vec -------
#if defined(__CUDA_ARCH__)
    vi_trp_pass(W, lo1, tid >> 5, tid & 31);
    VI_TRP_TICK(0)
    __syncthreads();
    VI_TRP_TICK(1)
#else
    for (int warp = 0; warp < (nt >> 5); ++warp) vi_trp_pass_host(W, lo1, warp);
#endif
    // ---- C: vector section (see vi_trp_c1 .. c3) ------------------------------------------------------------
#if defined(__CUDA_ARCH__)
    if (tid < nsub) {
      const bool in = (tid >= lo1) && (tid < n);
      const int warp = tid >> 5, lane = tid & 31, nw = nsub >> 5;
      double vn_i = 0.0, yv_i = 0.0, p_i = 0.0, r1 = 0.0, r2 = 0.0, xcol = 0.0;
      if (in) {
        vn_i = W.vn[tid]; yv_i = W.yv[tid];
        xcol = W.X[vi_trp_idx(npad, tid, lo1)];
        p_i = vi_trp_c1(W, lo1, tau, tid, vn_i, yv_i, &r1, &r2);
        W.p[tid] = p_i;
      }
      r1 = vi_bfly(r1); r2 = vi_bfly(r2);
      if (lane == 0) { W.part[2 * warp] = r1; W.part[2 * warp + 1] = r2; }
      VI_NAMED_BAR(nsub);
      VI_TRP_TICK(2)
      double dot = 0.0, dot2 = 0.0;
      for (int w = 0; w < nw; ++w) { dot += W.part[2 * w]; dot2 += W.part[2 * w + 1]; }
      double col_i = 0.0;
      if (in) col_i = vi_trp_c2(W, lo1, tau, dot, dot2, tid, p_i, vn_i, yv_i, xcol);
      if (k + 2 < n) {
        double sq = (tid >= k + 3 && tid < n) ? col_i * col_i : 0.0;
        sq = vi_bfly(sq);
        if (lane == 0) W.part2[warp] = sq;
        VI_NAMED_BAR(nsub);
        VI_TRP_TICK(3)
        double xn2 = 0.0;
        for (int w = 0; w < nw; ++w) xn2 += W.part2[w];
        vi_trp_c3(W, k + 1, xn2, V, tid, col_i);
      }
    }
    VI_TRP_TICK(4)
    __syncthreads();
    VI_TRP_TICK(5)
#else
    {
      const int nw = nsub >> 5;
      std::vector<double> vn_(nsub, 0.0), yv_(nsub, 0.0), p_(nsub, 0.0), r1_(nsub, 0.0), r2_(nsub, 0.0), xc_(nsub, 0.0),
          col_(nsub, 0.0), sq_(nsub, 0.0);
      for (int t = 0; t < nsub; ++t) {
        if (t >= lo1 && t < n) {
          vn_[t] = W.vn[t]; yv_[t] = W.yv[t];
          xc_[t] = W.X[vi_trp_idx(npad, t, lo1)];
          p_[t] = vi_trp_c1(W, lo1, tau, t, vn_[t], yv_[t], &r1_[t], &r2_[t]);
          W.p[t] = p_[t];
        }
      }
      for (int w = 0; w < nw; ++w) { vi_bfly_host(&r1_[32 * w], &W.part[2 * w]); vi_bfly_host(&r2_[32 * w], &W.part[2 * w + 1]); }
      double dot = 0.0, dot2 = 0.0;
      for (int w = 0; w < nw; ++w) { dot += W.part[2 * w]; dot2 += W.part[2 * w + 1]; }
      // (W.p[lo1] and W.vn[lo1] are read by every thread of C2 and written by none of them)
      for (int t = 0; t < nsub; ++t)
        if (t >= lo1 && t < n) col_[t] = vi_trp_c2(W, lo1, tau, dot, dot2, t, p_[t], vn_[t], yv_[t], xc_[t]);
      if (k + 2 < n) {
        for (int t = 0; t < nsub; ++t) sq_[t] = (t >= k + 3 && t < n) ? col_[t] * col_[t] : 0.0;
        for (int w = 0; w < nw; ++w) vi_bfly_host(&sq_[32 * w], &W.part2[w]);
        double xn2 = 0.0;
        for (int w = 0; w < nw; ++w) xn2 += W.part2[w];
        for (int t = 0; t < nsub; ++t) vi_trp_c3(W, k + 1, xn2, V, t, col_[t]);
      }
    }
#endif
  }
#if defined(__CUDA_ARCH__) && defined(VI_TRP_PROFILE)
  if ((tid == 0 || tid == 32 * 7) && (blockIdx.x % 2000) == 7)
    printf("[trp] block %d tid %d: pass %lld  bar1 %lld  C1 %lld  C2 %lld  C3 %lld  bar2 %lld  (cycles per system)\n",
           (int)blockIdx.x, tid, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5]);
#endif
  VI_PHASE(
    if (tid == 0) {
      if (n == 1) W.d[0] = W.X[0];
      else W.d[n - 1] = W.col[n - 1];
      W.e[n - 1] = 0.0; W.tau[n - 1] = 0.0;
    }
  )
}
