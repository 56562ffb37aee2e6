// Regularisation-parameter search state machines (one thread per record).
//
// Restates, as resumable state machines, two pieces of host control flow of the
// reference so that the expensive objective evaluations can be batched over all
// records between steps:
//   * Interpolate.chi2 (interpolate.py:173-218): scale-factor loop x decade walk
//     over alpha = 0,-1,...,-101.  f(alpha) = chi2(10^alpha) - nu does not depend
//     on the scale factor except through nu, so the walk is evaluated on a
//     precomputed table chi2[0..101] (bit-identical decisions, ~4x fewer solves).
//   * scipy.optimize.brentq as called at interpolate.py:214 (xtol = 2e-12,
//     rtol = 4*eps, maxiter = 100): Brent-Dekker with inverse quadratic / secant
//     steps guarded by bisection.
// VI_HD: unit-tested on the CPU against scipy (tests/cpu_harness.cpp).
#pragma once
#include <math.h>
#include <stdint.h>
#include "volinterp_b200.h"   // VI_NALPHA, VI_ST_* record status codes

#ifndef VI_HD
#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif
#endif

#define VI_BRENT_XTOL 2e-12
#define VI_BRENT_RTOL 8.881784197001252e-16
#define VI_BRENT_MAXITER 100

struct vi_bracket {
  int32_t status;      // VI_ST_OK => bracket valid
  int32_t k_lo;        // table index of alpha (lower end), alpha = -k_lo; alpha0 = -(k_lo-1)
  double nu;           // npts * scale factor that produced the bracket
};

// Decade walk of interpolate.py:180-207 on a table chi2[k] = chi2(10^-k).
VI_HD vi_bracket vi_chi2_bracket(const double* chi2, int64_t stride, int npts) {
  const double sfs[5] = {0.6, 0.7, 0.8, 0.9, 1.0};
  vi_bracket out;
  out.status = VI_ST_NO_ROOT;
  out.k_lo = -1;
  out.nu = 0.0;
  bool bracket = false;
  for (int isf = 0; isf < 5; ++isf) {
    double nu = npts * sfs[isf];
    double val0 = 1.0;
    int k = 0;
    double val = chi2[0] - nu;
    if (val < 0.0) { out.status = VI_ST_TOO_SMOOTH; out.nu = nu; return out; }
    while (val0 * val > 0.0) {
      bracket = true;
      val0 = val;
      k = k + 1;                       // alpha = alpha - 1
      val = chi2[(int64_t)k * stride] - nu;
      if (k > 100) { bracket = false; break; }   // alpha < -100
    }
    if (bracket) {
      out.status = VI_ST_OK;
      out.k_lo = k;
      out.nu = nu;
      return out;
    }
  }
  return out;
}

// Lazy table: with only chi2[0 .. avail) evaluated, would the walk above read an entry >= avail?  Only the
// first scale factor can be cut short: if its walk runs to k = 101 without a sign change the table is
// complete anyway and the later scale factors re-read it.  Same comparisons as vi_chi2_bracket.
VI_HD bool vi_chi2_walk_needs_more(const double* chi2, int64_t stride, int npts, int avail) {
  if (avail >= VI_NALPHA) return false;
  if (avail < 1) return true;
  const double nu = npts * 0.6;
  double val0 = 1.0;
  int k = 0;
  double val = chi2[0] - nu;
  if (val < 0.0) return false;
  while (val0 * val > 0.0) {
    val0 = val;
    k = k + 1;
    if (k >= avail) return true;
    val = chi2[(int64_t)k * stride] - nu;
    if (k > 100) break;
  }
  return false;
}

struct vi_brent {
  double xpre, xcur, xblk, fpre, fcur, fblk, spre, scur;
  double root;
  int32_t iter;
  int32_t done;   // 0 running, 1 converged, 2 iteration limit
};

// brentq(f, xa, xb) with f(xa), f(xb) known.  Returns true if already finished.
VI_HD bool vi_brent_init(vi_brent& b, double xa, double fa, double xb, double fb) {
  b.xpre = xa; b.xcur = xb; b.xblk = 0.0;
  b.fpre = fa; b.fcur = fb; b.fblk = 0.0;
  b.spre = 0.0; b.scur = 0.0;
  b.iter = 0; b.done = 0; b.root = xb;
  if (fa == 0.0) { b.root = xa; b.done = 1; }
  else if (fb == 0.0) { b.root = xb; b.done = 1; }
  return b.done != 0;
}

// Advance to the next abscissa to evaluate (returned in b.xcur).  Returns true
// when the search has terminated (b.root valid, b.done set).
VI_HD bool vi_brent_propose(vi_brent& b) {
  if (b.done) return true;
  if (b.iter >= VI_BRENT_MAXITER) { b.done = 2; b.root = b.xcur; return true; }
  if (b.fpre != 0.0 && b.fcur != 0.0 && (signbit(b.fpre) != signbit(b.fcur))) {
    b.xblk = b.xpre; b.fblk = b.fpre;
    b.spre = b.scur = b.xcur - b.xpre;
  }
  if (fabs(b.fblk) < fabs(b.fcur)) {
    b.xpre = b.xcur; b.xcur = b.xblk; b.xblk = b.xpre;
    b.fpre = b.fcur; b.fcur = b.fblk; b.fblk = b.fpre;
  }
  double delta = (VI_BRENT_XTOL + VI_BRENT_RTOL * fabs(b.xcur)) / 2.0;
  double sbis = (b.xblk - b.xcur) / 2.0;
  if (b.fcur == 0.0 || fabs(sbis) < delta) { b.root = b.xcur; b.done = 1; return true; }
  if (fabs(b.spre) > delta && fabs(b.fcur) < fabs(b.fpre)) {
    double stry;
    if (b.xpre == b.xblk) {
      stry = -b.fcur * (b.xcur - b.xpre) / (b.fcur - b.fpre);
    } else {
      double dpre = (b.fpre - b.fcur) / (b.xpre - b.xcur);
      double dblk = (b.fblk - b.fcur) / (b.xblk - b.xcur);
      stry = -b.fcur * (b.fblk * dblk - b.fpre * dpre) / (dblk * dpre * (b.fblk - b.fpre));
    }
    double lim = fmin(fabs(b.spre), 3.0 * fabs(sbis) - delta);
    if (2.0 * fabs(stry) < lim) { b.spre = b.scur; b.scur = stry; }
    else { b.spre = sbis; b.scur = sbis; }
  } else {
    b.spre = sbis; b.scur = sbis;
  }
  b.xpre = b.xcur; b.fpre = b.fcur;
  if (fabs(b.scur) > delta) b.xcur += b.scur;
  else b.xcur += (sbis > 0.0 ? delta : -delta);
  return false;
}

VI_HD void vi_brent_feed(vi_brent& b, double fnew) {
  b.fcur = fnew;
  b.iter += 1;
}
