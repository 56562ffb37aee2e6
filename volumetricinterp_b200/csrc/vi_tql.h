// Symmetric tridiagonal eigen-solve by implicit-shift QL with a ROTATION TAPE.
//
// Role on the hot path (kernel K3b): the reference solves
// (A^T W A + lambda*R) C = A^T W d with scipy.linalg.lstsq == LAPACK gelsd with
// rcond = eps (interpolate.py:462), i.e. the minimum-norm solution over the
// singular values s_i > eps*s_max.  For the (symmetrised) matrix that is
//   C = sum_{|l_i| > eps*max|l|}  v_i (v_i . y) / l_i .
// The CTA kernel reduces X to tridiagonal T = Q^T X Q; this routine (one THREAD
// per system, thousands of systems in flight) diagonalises T = Z L Z^T without
// ever forming Z: every Givens rotation is applied on the fly to g = Z^T y'
// and appended to a tape; Z*u is obtained afterwards by replaying the tape
// backwards.  Cost O(n^2) per system instead of O(n^3).
//
// Like LAPACK dsteqr, each unreduced block is processed from its smaller-|d| end
// (on the graded matrices this path produces, deflating from the large end
// loses the small eigenvalues and changes the numerical rank — see DESIGN.md).
//
// VI_HD: compiled for the device in fit.cu and for the CPU in the test-only
// harness (tests/cpu_harness.cpp) where it is checked against LAPACK.
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

#define VI_EPS_HALF 1.1102230246251565e-16   // 2^-53 (LAPACK dlamch('E'))
#define VI_EPS 2.220446049250313e-16         // numpy finfo(float64).eps, scipy lstsq rcond
#define VI_SAFMIN 2.2250738585072014e-308

// Strided view: element i of the vector owned by one thread lives at p[i*stride].
// stride = 32 gives the warp-interleaved layout used on the GPU (coalesced when the
// lanes of a warp index the same i); stride = 1 on the CPU.
struct vi_svec {
  double* p;
  int64_t stride;
  VI_HD double& operator[](int64_t i) const { return p[i * stride]; }
};
struct vi_sidx {
  int32_t* p;
  int64_t stride;
  VI_HD int32_t& operator[](int64_t i) const { return p[i * stride]; }
};

struct vi_tape {
  vi_svec c, s;     // rotation cosine / sine
  vi_sidx ix;       // (physical index of logical i) * 2 + (1 if partner is index-1 else 0)
  int32_t cap;      // capacity in rotations
};

VI_HD double vi_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// status: 0 ok, 1 iteration limit, 2 tape overflow
VI_HD int vi_tql(int n, vi_svec d, vi_svec e, vi_svec g, vi_tape tape, int32_t* nrot_out) {
  int32_t nrot = 0;
  int status = 0;
  int budget = 30 * n;                     // LAPACK's nmaxit
  int l1 = 0;
  while (l1 < n) {
    // ---- delimit the next unreduced block [l1, m] ----
    int m = l1;
    while (m < n - 1) {
      double tst = fabs(e[m]);
      if (tst == 0.0) break;
      if (tst <= (sqrt(fabs(d[m])) * sqrt(fabs(d[m + 1]))) * VI_EPS_HALF) { e[m] = 0.0; break; }
      ++m;
    }
    const int lo = l1, hi = m;
    l1 = m + 1;
    const int nb = hi - lo + 1;
    if (nb == 1) continue;
    // ---- exact power-of-two scaling of the block ----
    double anorm = 0.0;
    for (int i = lo; i <= hi; ++i) {
      anorm = fmax(anorm, fabs(d[i]));
      if (i < hi) anorm = fmax(anorm, fabs(e[i]));
    }
    if (anorm == 0.0) continue;
    int ex;
    frexp(anorm, &ex);
    const double scl = ldexp(1.0, -ex), uns = ldexp(1.0, ex);
    for (int i = lo; i <= hi; ++i) {
      d[i] = d[i] * scl;
      if (i < hi) e[i] = e[i] * scl;
    }
    // ---- logical view: QL deflates at logical index 0, which must be the small end ----
    const bool rev = fabs(d[hi]) < fabs(d[lo]);
#define PD(i) (rev ? (hi - (i)) : (lo + (i)))
#define PE(i) (rev ? (hi - (i)-1) : (lo + (i)))
    for (int l = 0; l < nb; ++l) {
      for (;;) {
        int mm = l;
        while (mm < nb - 1) {
          double ev = e[PE(mm)];
          double tst = ev * ev;
          if (tst <= (VI_EPS_HALF * VI_EPS_HALF * fabs(d[PD(mm)])) * fabs(d[PD(mm + 1)]) + VI_SAFMIN) break;
          ++mm;
        }
        if (mm < nb - 1) e[PE(mm)] = 0.0;
        if (mm == l) break;
        if (budget-- <= 0) { status = 1; goto done_block; }
        // Wilkinson shift from the leading 2x2 of the active block
        double el = e[PE(l)];
        double gg = (d[PD(l + 1)] - d[PD(l)]) / (2.0 * el);
        double r = hypot(gg, 1.0);
        gg = d[PD(mm)] - d[PD(l)] + el / (gg + vi_sign(r, gg));
        double s = 1.0, c = 1.0, p = 0.0;
        int i = mm - 1;
        bool early = false;
        for (; i >= l; --i) {
          double ei = e[PE(i)];
          double f = s * ei, b = c * ei;
          // the block is scaled to max|.| ~ 1, so f*f + gg*gg cannot overflow; squares that underflow are
          // below 1e-308 relative to 1 and are treated as zero (the r == 0 branch, as in LAPACK)
          const double r2 = f * f + gg * gg;
          if (r2 == 0.0) {
            if (i + 1 < nb - 1) e[PE(i + 1)] = 0.0;
            d[PD(i + 1)] -= p;
            if (mm < nb - 1) e[PE(mm)] = 0.0;
            early = true;
            break;
          }
          // one reciprocal square root instead of hypot + two divisions (the dependent chain of every
          // rotation runs through here): r = r2 / sqrt(r2), s = f / r, c = g / r
#if defined(__CUDA_ARCH__)
          const double ri = rsqrt(r2);
#else
          const double ri = 1.0 / sqrt(r2);
#endif
          r = r2 * ri;
          if (i + 1 < nb - 1) e[PE(i + 1)] = r;
          s = f * ri;
          c = gg * ri;
          gg = d[PD(i + 1)] - p;
          r = (d[PD(i)] - gg) * s + 2.0 * c * b;
          p = s * r;
          d[PD(i + 1)] = gg + p;
          gg = c * r - b;
          // apply to g (row vector times rotation in the logical plane (i, i+1))
          const int pi = PD(i), pj = PD(i + 1);
          double gi = g[pi], gj = g[pj];
          g[pj] = s * gi + c * gj;
          g[pi] = c * gi - s * gj;
          if (nrot < tape.cap) {
            tape.c[nrot] = c;
            tape.s[nrot] = s;
            tape.ix[nrot] = pi * 2 + (rev ? 1 : 0);
          } else {
            status = 2;
          }
          ++nrot;
        }
        if (early) continue;
        d[PD(l)] -= p;
        e[PE(l)] = gg;
        if (mm < nb - 1) e[PE(mm)] = 0.0;
      }
    }
  done_block:
#undef PD
#undef PE
    for (int i = lo; i <= hi; ++i) d[i] = d[i] * uns;
    if (status == 1) break;
  }
  *nrot_out = nrot;
  return status;
}

// Eigenvalues + rotation tape only (no right-hand side): the variant the GPU hot path runs, one
// thread per system with d and e in shared memory.  Same algorithm and deflation rules as vi_tql
// above; the differences are mechanical, to shorten the dependent chain and the instruction count of
// the inner loop:
//   * a block whose small end is at the top is physically reversed first (and reversed back at the
//     end), so the loops index d/e directly instead of through the PD/PE view;
//   * running pointers instead of index arithmetic; (c, s) stored as one pair;
//   * Z^T g is NOT accumulated here - the tape is replayed forwards by the apply kernel.
// tape.c / tape.s must be the two halves of interleaved (c, s) pairs (c.p + 1 == s.p, stride 2).
VI_HD int vi_tql_values(int n, vi_svec d, vi_svec e, vi_tape tape, int32_t* nrot_out) {
  int32_t nrot = 0;
  int status = 0;
  int budget = 30 * n;
  const int64_t sd = d.stride, se = e.stride;
  double* const cs = tape.c.p;        // pairs: cs[2t] = c, cs[2t+1] = s
  int32_t* const ix = tape.ix.p;
  const int32_t cap = tape.cap;
  int l1 = 0;
  while (l1 < n) {
    int m = l1;
    while (m < n - 1) {
      double tst = fabs(e[m]);
      if (tst == 0.0) break;
      if (tst <= (sqrt(fabs(d[m])) * sqrt(fabs(d[m + 1]))) * VI_EPS_HALF) { e[m] = 0.0; break; }
      ++m;
    }
    const int lo = l1, hi = m;
    l1 = m + 1;
    const int nb = hi - lo + 1;
    if (nb == 1) continue;
    double anorm = 0.0;
    for (int i = lo; i <= hi; ++i) {
      anorm = fmax(anorm, fabs(d[i]));
      if (i < hi) anorm = fmax(anorm, fabs(e[i]));
    }
    if (anorm == 0.0) continue;
    int ex;
    frexp(anorm, &ex);
    const double scl = ldexp(1.0, -ex), uns = ldexp(1.0, ex);
    const bool rev = fabs(d[hi]) < fabs(d[lo]);
    // scale, and lay the block out so that logical index 0 (where QL deflates) is the small end
    if (!rev) {
      for (int i = lo; i <= hi; ++i) { d[i] = d[i] * scl; if (i < hi) e[i] = e[i] * scl; }
    } else {
      for (int a = lo, b = hi; a <= b; ++a, --b) {
        double da = d[a] * scl, db = d[b] * scl;
        d[a] = db; if (a != b) d[b] = da;
      }
      for (int a = lo, b = hi - 1; a <= b; ++a, --b) {
        double ea = e[a] * scl, eb = e[b] * scl;
        e[a] = eb; if (a != b) e[b] = ea;
      }
    }
    double* const D = d.p + (int64_t)lo * sd;     // D[i*sd], E[i*se]: logical i = 0..nb-1
    double* const E = e.p + (int64_t)lo * se;
    // physical index of logical i: lo + i (not reversed) or hi - i (reversed); partner = logical i+1
    const int pbase = rev ? hi : lo;
    const int pstep = rev ? -1 : 1;
    for (int l = 0; l < nb; ++l) {
      for (;;) {
        int mm = l;
        {
          const double* dp = D + (int64_t)l * sd;
          const double* ep = E + (int64_t)l * se;
          double dcur = fabs(*dp);
          while (mm < nb - 1) {
            const double ev = *ep;
            const double dnx = fabs(dp[sd]);
            if (ev * ev <= (VI_EPS_HALF * VI_EPS_HALF * dcur) * dnx + VI_SAFMIN) break;
            dcur = dnx; dp += sd; ep += se; ++mm;
          }
        }
        if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
        if (mm == l) break;
        if (budget-- <= 0) { status = 1; goto done_block; }
        const double el = E[(int64_t)l * se];
        const double dl = D[(int64_t)l * sd];
        double gg = (D[(int64_t)(l + 1) * sd] - dl) / (2.0 * el);
        double r = sqrt(gg * gg + 1.0);
        gg = D[(int64_t)mm * sd] - dl + el / (gg + vi_sign(r, gg));
        double s = 1.0, c = 1.0, p = 0.0;
        int i = mm - 1;
        double* dp = D + (int64_t)i * sd;          // d[i]; dp[sd] = d[i+1]
        double* ep = E + (int64_t)i * se;          // e[i]; ep[se] = e[i+1]
        double dnext = dp[sd];                     // d[i+1], carried in a register between steps
        // e[i] and d[i] of the NEXT step are loaded one step ahead (this step writes e[i+1] and d[i+1] only): two
        // shared-memory latencies less on the dependent chain of every rotation
        double e_pre = *ep, d_pre = *dp;
        bool early = false;
        for (; i >= l; --i) {
          const double ei = e_pre;
          const double di_pre = d_pre;
          if (i > l) { e_pre = ep[-se]; d_pre = dp[-sd]; }
          const double f = s * ei, b = c * ei;
          const double r2 = f * f + gg * gg;
          if (r2 == 0.0) {
            if (i + 1 < nb - 1) ep[se] = 0.0;
            dp[sd] = dnext - p;
            if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
            early = true;
            break;
          }
#if defined(__CUDA_ARCH__)
          const double ri = rsqrt(r2);
#else
          const double ri = 1.0 / sqrt(r2);
#endif
          r = r2 * ri;
          if (i + 1 < nb - 1) ep[se] = r;
          s = f * ri;
          c = gg * ri;
          gg = dnext - p;
          const double di = di_pre;
          r = (di - gg) * s + 2.0 * c * b;
          p = s * r;
          dp[sd] = gg + p;
          gg = c * r - b;
          dnext = di;
          if (nrot < cap) {
            cs[2 * (int64_t)nrot] = c;
            cs[2 * (int64_t)nrot + 1] = s;
            ix[nrot] = (pbase + pstep * i) * 2 + (rev ? 1 : 0);
          } else {
            status = 2;
          }
          ++nrot;
          dp -= sd; ep -= se;
        }
        if (early) continue;
        D[(int64_t)l * sd] = dnext - p;
        E[(int64_t)l * se] = gg;
        if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
      }
    }
  done_block:
    // unscale and restore the physical order
    if (!rev) {
      for (int i = lo; i <= hi; ++i) d[i] = d[i] * uns;
    } else {
      for (int a = lo, b = hi; a <= b; ++a, --b) {
        double da = d[a] * uns, db = d[b] * uns;
        d[a] = db; if (a != b) d[b] = da;
      }
    }
    if (status == 1) break;
  }
  *nrot_out = nrot;
  return status;
}

// Same algorithm as vi_tql_values, FLATTENED into a per-lane state machine so that the 32 independent
// systems of a warp spend their time in one common loop body (one rotation per trip) instead of
// serialising nested loops of different trip counts (measured: 10 of 32 lanes active in the nested form).
// Two changes of organisation, none of arithmetic (eigenvalues and tape are bit-identical, see
// tests/test_host_math.py):
//   * the search for the first negligible off-diagonal element is fused into the sweep: every e[m] is
//     tested right after the sweep has produced its final value, so the separate O(n) scan per QL
//     iteration is only needed at the start of a block, after a zero rotation, or after two eigenvalues
//     converged in one sweep;
//   * block set-up / tear-down (rare, O(n)) are phases of the same machine.
// `active` = false makes the lane idle (device: every lane of the warp must call this function).
enum { VI_QL_BLOCK = 0, VI_QL_ITER, VI_QL_ROT, VI_QL_UNBLOCK, VI_QL_DONE };

VI_HD bool vi_ql_negl(double ev, double da, double db) {
  return ev * ev <= (VI_EPS_HALF * VI_EPS_HALF * fabs(da)) * fabs(db) + VI_SAFMIN;
}

VI_HD int vi_tql_values_flat(int n, vi_svec d, vi_svec e, vi_tape tape, int32_t* nrot_out, bool active,
                              int iter_batch = 8) {
  int32_t nrot = 0;
  int status = 0;
  int budget = 30 * n;
  const int64_t sd = d.stride, se = e.stride;
  double* const cs = tape.c.p;
  int32_t* const ix = tape.ix.p;
  const int32_t cap = tape.cap;
  int phase = active ? VI_QL_BLOCK : VI_QL_DONE;
  // block state
  int l1 = 0, lo = 0, hi = 0, nb = 0, pbase = 0, pstep = 1;
  bool rev = false;
  double uns = 1.0;
  double* D = d.p;
  double* E = e.p;
  // iteration state
  int l = 0, mm = 0, cand = 0, mt1 = -1, i = 0;
  bool valid = false, cand_valid = false;
  double s = 1.0, c = 1.0, p = 0.0, gg = 0.0, dnext = 0.0;
  double e_pre = 0.0, d_pre = 0.0;       // e[i], d[i] of the next rotation, loaded one trip ahead (see vi_tql_values)
  for (;;) {
    // The sweep set-up (shift: two divisions and a square root) is batched: lanes that reach it idle until
    // `iter_batch` of them are waiting or no lane is rotating, so its ~120 instructions are not paid on every trip.
#if defined(__CUDA_ARCH__)
    const unsigned m_iter = __ballot_sync(0xffffffffu, phase == VI_QL_ITER);
    const unsigned m_rot = __ballot_sync(0xffffffffu, phase == VI_QL_ROT);
    const bool do_iter = (__popc(m_iter) >= iter_batch) || (m_rot == 0u);
#else
    const bool do_iter = true;
#endif
    if (phase == VI_QL_ROT) {
      double* dp = D + (int64_t)i * sd;
      double* ep = E + (int64_t)i * se;
      const double ei = e_pre;
      const double di_pre = d_pre;
      if (i > l) { e_pre = ep[-se]; d_pre = dp[-sd]; }
      const double f = s * ei, b = c * ei;
      const double r2 = f * f + gg * gg;
      if (r2 == 0.0) {
        if (i + 1 < nb - 1) ep[se] = 0.0;
        dp[sd] = dnext - p;
        if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
        valid = false;
        phase = VI_QL_ITER;
      } else {
#if defined(__CUDA_ARCH__)
        const double ri = rsqrt(r2);
#else
        const double ri = 1.0 / sqrt(r2);
#endif
        double r = r2 * ri;
        if (i + 1 < nb - 1) ep[se] = r;
        s = f * ri;
        c = gg * ri;
        gg = dnext - p;
        const double di = di_pre;
        r = (di - gg) * s + 2.0 * c * b;
        p = s * r;
        const double dnew = gg + p;          // final d[i+1] of this sweep
        dp[sd] = dnew;
        gg = c * r - b;
        // fused scan: e[i+1] (= the value just stored) is final; d[i+2] was finalised one step earlier
        if (i + 1 <= mm - 1 && vi_ql_negl(r2 * ri, dnew, D[(int64_t)(i + 2) * sd])) mt1 = i + 1;
        dnext = di;
        if (nrot < cap) {
          cs[2 * (int64_t)nrot] = c;
          cs[2 * (int64_t)nrot + 1] = s;
          ix[nrot] = (pbase + pstep * i) * 2 + (rev ? 1 : 0);
        } else {
          status = 2;
        }
        ++nrot;
        --i;
        if (i < l) {
          // end of the sweep
          const double dl = dnext - p;
          D[(int64_t)l * sd] = dl;
          E[(int64_t)l * se] = gg;
          if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
          const int above = (mt1 >= 0) ? mt1 : mm;       // first negligible index in [l+1, mm]
          if (vi_ql_negl(gg, dl, D[(int64_t)(l + 1) * sd])) { mm = l; cand = above; cand_valid = true; }
          else { mm = above; cand_valid = false; }
          valid = true;
          phase = VI_QL_ITER;
        }
      }
    } else if (phase == VI_QL_ITER && do_iter) {
      for (;;) {
        if (!valid) {
          mm = l;
          while (mm < nb - 1 && !vi_ql_negl(E[(int64_t)mm * se], D[(int64_t)mm * sd], D[(int64_t)(mm + 1) * sd])) ++mm;
          valid = true;
          cand_valid = false;
        }
        if (mm < nb - 1) E[(int64_t)mm * se] = 0.0;
        if (mm == l) {
          ++l;
          if (l >= nb) { phase = VI_QL_UNBLOCK; break; }
          if (cand_valid && cand >= l) {
            mm = cand;              // known from the last sweep; anything beyond it is not
            cand_valid = false;
          } else {
            valid = false;
          }
          continue;
        }
        if (budget-- <= 0) { status = 1; phase = VI_QL_UNBLOCK; break; }
        const double el = E[(int64_t)l * se];
        const double dl = D[(int64_t)l * sd];
        double g0 = (D[(int64_t)(l + 1) * sd] - dl) / (2.0 * el);
        const double r = sqrt(g0 * g0 + 1.0);
        gg = D[(int64_t)mm * sd] - dl + el / (g0 + vi_sign(r, g0));
        s = 1.0; c = 1.0; p = 0.0;
        i = mm - 1;
        dnext = D[(int64_t)(i + 1) * sd];
        e_pre = E[(int64_t)i * se];
        d_pre = D[(int64_t)i * sd];
        mt1 = -1;
        phase = VI_QL_ROT;
        break;
      }
    } else if (phase == VI_QL_BLOCK) {
      if (l1 >= n) {
        phase = VI_QL_DONE;
      } else {
        int m = l1;
        while (m < n - 1) {
          double tst = fabs(e[m]);
          if (tst == 0.0) break;
          if (tst <= (sqrt(fabs(d[m])) * sqrt(fabs(d[m + 1]))) * VI_EPS_HALF) { e[m] = 0.0; break; }
          ++m;
        }
        lo = l1; hi = m; l1 = m + 1; nb = hi - lo + 1;
        double anorm = 0.0;
        for (int k = lo; k <= hi; ++k) {
          anorm = fmax(anorm, fabs(d[k]));
          if (k < hi) anorm = fmax(anorm, fabs(e[k]));
        }
        if (nb > 1 && anorm != 0.0) {
          int ex;
          frexp(anorm, &ex);
          const double scl = ldexp(1.0, -ex);
          uns = ldexp(1.0, ex);
          rev = fabs(d[hi]) < fabs(d[lo]);
          if (!rev) {
            for (int k = lo; k <= hi; ++k) { d[k] = d[k] * scl; if (k < hi) e[k] = e[k] * scl; }
          } else {
            for (int a = lo, b = hi; a <= b; ++a, --b) {
              double da = d[a] * scl, db = d[b] * scl;
              d[a] = db; if (a != b) d[b] = da;
            }
            for (int a = lo, b = hi - 1; a <= b; ++a, --b) {
              double ea = e[a] * scl, eb = e[b] * scl;
              e[a] = eb; if (a != b) e[b] = ea;
            }
          }
          D = d.p + (int64_t)lo * sd;
          E = e.p + (int64_t)lo * se;
          pbase = rev ? hi : lo;
          pstep = rev ? -1 : 1;
          l = 0; valid = false; cand_valid = false;
          phase = VI_QL_ITER;
        }
      }
    } else if (phase == VI_QL_UNBLOCK) {
      if (!rev) {
        for (int k = lo; k <= hi; ++k) d[k] = d[k] * uns;
      } else {
        for (int a = lo, b = hi; a <= b; ++a, --b) {
          double da = d[a] * uns, db = d[b] * uns;
          d[a] = db; if (a != b) d[b] = da;
        }
      }
      phase = (status == 1) ? VI_QL_DONE : VI_QL_BLOCK;
    }
#if defined(__CUDA_ARCH__)
    if (__all_sync(0xffffffffu, phase == VI_QL_DONE)) break;
#else
    if (phase == VI_QL_DONE) break;
#endif
  }
  *nrot_out = nrot;
  return status;
}

// w <- Z w  (replay the tape backwards).  The tape lives in global memory: four entries are fetched
// ahead of the (dependent) updates of w so their latency overlaps.
VI_HD void vi_tape_apply_z(vi_svec w, vi_tape tape, int32_t nrot) {
  int32_t t = nrot - 1;
  for (; t >= 3; t -= 4) {
    int32_t code[4]; double c[4], s[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) { code[q] = tape.ix[t - q]; c[q] = tape.c[t - q]; s[q] = tape.s[t - q]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      int pi = code[q] >> 1;
      int pj = (code[q] & 1) ? pi - 1 : pi + 1;
      double a = w[pi], b = w[pj];
      w[pi] = c[q] * a + s[q] * b;
      w[pj] = c[q] * b - s[q] * a;
    }
  }
  for (; t >= 0; --t) {
    int32_t code = tape.ix[t];
    int pi = code >> 1;
    int pj = (code & 1) ? pi - 1 : pi + 1;
    double c = tape.c[t], s = tape.s[t];
    double a = w[pi], b = w[pj];
    w[pi] = c * a + s * b;
    w[pj] = c * b - s * a;
  }
}

// g <- Z^T g  (replay the tape forwards), four entries fetched ahead.
VI_HD void vi_tape_apply_zt(vi_svec g, vi_tape tape, int32_t nrot) {
  int32_t t = 0;
  for (; t + 3 < nrot; t += 4) {
    int32_t code[4]; double c[4], s[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) { code[q] = tape.ix[t + q]; c[q] = tape.c[t + q]; s[q] = tape.s[t + q]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      int pi = code[q] >> 1;
      int pj = (code[q] & 1) ? pi - 1 : pi + 1;
      double gi = g[pi], gj = g[pj];
      g[pj] = s[q] * gi + c[q] * gj;
      g[pi] = c[q] * gi - s[q] * gj;
    }
  }
  for (; t < nrot; ++t) {
    int32_t code = tape.ix[t];
    int pi = code >> 1;
    int pj = (code & 1) ? pi - 1 : pi + 1;
    double c = tape.c[t], s = tape.s[t];
    double gi = g[pi], gj = g[pj];
    g[pj] = s * gi + c * gj;
    g[pi] = c * gi - s * gj;
  }
}

// u_i = g_i / l_i over |l_i| > rcond * max|l|, else 0 (gelsd / pinv cut-off).  Returns rank.
VI_HD int vi_spectral_divide(int n, vi_svec lam, vi_svec g, double rcond) {
  double lmax = 0.0;
  for (int i = 0; i < n; ++i) lmax = fmax(lmax, fabs(lam[i]));
  const double cut = rcond * lmax;
  int rank = 0;
  for (int i = 0; i < n; ++i) {
    if (fabs(lam[i]) > cut) { g[i] = g[i] / lam[i]; ++rank; }
    else g[i] = 0.0;
  }
  return rank;
}
