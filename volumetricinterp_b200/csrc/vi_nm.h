// One-dimensional Nelder-Mead as a resumable state machine (one thread per search unit).
//
// Restates scipy.optimize.minimize(method='Nelder-Mead') exactly as the reference calls it for the GCV
// regularisation parameter (interpolate.py:288-291: x0 = -20, every option at its default): simplex
// {x0, 1.05 x0}, rho = 1, chi = 2, psi = 0.5, sigma = 0.5, xatol = fatol = 1e-4, maxiter = maxfev = 200,
// including the order of the comparisons (which decides what happens with NaN objective values), the
// evaluation-count guard (an evaluation attempted when maxfev is reached aborts the iteration) and the
// success rule (failure if either limit was hit -> the reference raises ValueError -> NaN parameter,
// interpolate.py:292-293, 142-145).  The objective is evaluated OUTSIDE (a batch of eigen-systems per
// abscissa), hence the next()/feed() split.  VI_HD: unit-tested on the CPU against scipy.
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

#define VI_NM_XATOL 1e-4
#define VI_NM_FATOL 1e-4
#define VI_NM_MAXITER 200
#define VI_NM_MAXFUN 200

enum { VI_NM_INIT0 = 0, VI_NM_INIT1, VI_NM_TOP, VI_NM_REQ, VI_NM_WAIT_R, VI_NM_WAIT_E, VI_NM_WAIT_C, VI_NM_WAIT_CC,
       VI_NM_WAIT_S, VI_NM_DONE };

struct vi_nm {
  double x0, x1, f0, f1;     // simplex, sorted so that f0 <= f1 (NaN last)
  double xr, fxr, xt;        // reflected point, its value, pending trial abscissa
  int32_t phase, next_phase, iters, fcalls, success;
};

// products / sums that must not be contracted into FMAs (scipy evaluates them as separate numpy ops)
VI_HD double vi_nm_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
VI_HD double vi_nm_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

VI_HD void vi_nm_sort(vi_nm& s) {
  // np.argsort on two values: ascending, NaN last, ties keep their order
  const bool swap = (s.f1 < s.f0) || ((s.f0 != s.f0) && (s.f1 == s.f1));
  if (swap) { double t = s.x0; s.x0 = s.x1; s.x1 = t; t = s.f0; s.f0 = s.f1; s.f1 = t; }
}

VI_HD void vi_nm_init(vi_nm& s, double x0) {
  s.x0 = x0;
  s.x1 = (x0 != 0.0) ? vi_nm_mul(1.05, x0) : 0.00025;     // (1 + nonzdelt) * x0 with 1 + 0.05 == 1.05 in binary64
  s.f0 = s.f1 = INFINITY;
  s.xr = s.fxr = s.xt = 0.0;
  s.phase = VI_NM_INIT0;
  s.next_phase = 0;
  s.iters = 0;
  s.fcalls = 0;
  s.success = 0;
}

VI_HD void vi_nm_finish(vi_nm& s) {
  s.success = (s.fcalls >= VI_NM_MAXFUN || s.iters >= VI_NM_MAXITER) ? 0 : 1;
  s.phase = VI_NM_DONE;
}

// Abscissa of the next objective evaluation in *x; false when the search has terminated (s.x0 is the
// minimiser, s.success says whether scipy would have reported success).
VI_HD bool vi_nm_next(vi_nm& s, double* x) {
  for (;;) {
    switch (s.phase) {
      case VI_NM_INIT0: s.xt = s.x0; s.next_phase = VI_NM_INIT0; s.phase = VI_NM_REQ; break;
      case VI_NM_INIT1: s.xt = s.x1; s.next_phase = VI_NM_INIT1; s.phase = VI_NM_REQ; break;
      case VI_NM_TOP: {
        if (!(s.fcalls < VI_NM_MAXFUN && s.iters < VI_NM_MAXITER)) { vi_nm_finish(s); return false; }
        if (fabs(s.x1 - s.x0) <= VI_NM_XATOL && fabs(s.f0 - s.f1) <= VI_NM_FATOL) { vi_nm_finish(s); return false; }
        s.xr = vi_nm_add(vi_nm_mul(2.0, s.x0), -vi_nm_mul(1.0, s.x1));     // (1 + rho) xbar - rho sim[-1]
        s.xt = s.xr; s.next_phase = VI_NM_WAIT_R; s.phase = VI_NM_REQ;
        break;
      }
      case VI_NM_REQ: {
        if (s.fcalls >= VI_NM_MAXFUN) {
          // _MaxFuncCallError: the rest of the iteration is skipped (for the two initial evaluations the
          // value stays +inf), the simplex is re-sorted and the loop condition ends the search
          vi_nm_sort(s);
          s.phase = VI_NM_TOP;
          break;
        }
        s.fcalls += 1;
        *x = s.xt;
        s.phase = s.next_phase;
        return true;
      }
      case VI_NM_DONE: return false;
      default: return false;     // waiting for a value: caller error
    }
  }
}

VI_HD void vi_nm_feed(vi_nm& s, double f) {
  switch (s.phase) {
    case VI_NM_INIT0: s.f0 = f; s.phase = VI_NM_INIT1; break;
    case VI_NM_INIT1: s.f1 = f; vi_nm_sort(s); vi_nm_sort(s); s.iters = 1; s.phase = VI_NM_TOP; break;
    case VI_NM_WAIT_R:
      s.fxr = f;
      if (s.fxr < s.f0) {
        s.xt = vi_nm_add(vi_nm_mul(3.0, s.x0), -vi_nm_mul(2.0, s.x1));      // (1 + rho chi) xbar - rho chi sim[-1]
        s.next_phase = VI_NM_WAIT_E; s.phase = VI_NM_REQ;
      } else if (s.fxr < s.f0) {            // fxr < fsim[-2]: fsim[-2] is fsim[0] in one dimension, never true here
        s.x1 = s.xr; s.f1 = s.fxr; s.iters += 1; vi_nm_sort(s); s.phase = VI_NM_TOP;
      } else if (s.fxr < s.f1) {
        s.xt = vi_nm_add(vi_nm_mul(1.5, s.x0), -vi_nm_mul(0.5, s.x1));      // (1 + psi rho) xbar - psi rho sim[-1]
        s.next_phase = VI_NM_WAIT_C; s.phase = VI_NM_REQ;
      } else {
        s.xt = vi_nm_add(vi_nm_mul(0.5, s.x0), vi_nm_mul(0.5, s.x1));       // (1 - psi) xbar + psi sim[-1]
        s.next_phase = VI_NM_WAIT_CC; s.phase = VI_NM_REQ;
      }
      break;
    case VI_NM_WAIT_E:
      if (f < s.fxr) { s.x1 = s.xt; s.f1 = f; } else { s.x1 = s.xr; s.f1 = s.fxr; }
      s.iters += 1; vi_nm_sort(s); s.phase = VI_NM_TOP;
      break;
    case VI_NM_WAIT_C:
    case VI_NM_WAIT_CC: {
      const bool accept = (s.phase == VI_NM_WAIT_C) ? (f <= s.fxr) : (f < s.f1);
      if (accept) {
        s.x1 = s.xt; s.f1 = f; s.iters += 1; vi_nm_sort(s); s.phase = VI_NM_TOP;
      } else {
        s.x1 = vi_nm_add(s.x0, vi_nm_mul(0.5, vi_nm_add(s.x1, -s.x0)));     // sim[0] + sigma (sim[j] - sim[0])
        s.xt = s.x1; s.next_phase = VI_NM_WAIT_S; s.phase = VI_NM_REQ;
      }
      break;
    }
    case VI_NM_WAIT_S: s.f1 = f; s.iters += 1; vi_nm_sort(s); s.phase = VI_NM_TOP; break;
    default: break;
  }
}
