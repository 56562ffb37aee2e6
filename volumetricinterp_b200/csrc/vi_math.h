// Per-point scalar mathematics of the two model plug-ins (sphharmlag, radbasfun).
//
// Every routine is VI_HD (host+device) and free of CUDA-only constructs so the
// same source is unit-tested on the CPU against scipy (tests/cpu_harness.cpp,
// test-only) and compiled into the sm_100a kernels (basis.cu, estimate.cu).
// All arithmetic is IEEE FP64; fused contraction is disabled at build time
// (-fmad=false) for the translation units that include this header so the
// operation order below is what executes.
//
// Reference behaviour restated (file:line under /root/reference/volumetricinterp):
//   geodetic -> ECEF            pymap3d.geodetic2ecef (models/sphharmlag.py:345,351)
//   model coordinates           models/sphharmlag.py:345-359
//   Laguerre L_k                scipy.special.eval_laguerre (models/sphharmlag.py:141)
//   Legendre P_v^m, real v      scipy.special.lpmv = Zhang & Jin LPMV/LPMV0 (sphharmlag.py:141)
//   azimuth A_vm, K_vm          models/sphharmlag.py:278-281, 318-321
//   gradient of the basis       models/sphharmlag.py:148-184 (dAz :298-301, scipy eval_genlaguerre)
//   Gaussian RBF                models/radbasfun.py:102-107
#pragma once
#include <math.h>
#include <stdint.h>
#include "volinterp_b200.h"   // vi_shl_params, VI_MAXL_MAX, VI_MAXK_MAX

#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif

#define VI_WGS84_A 6378137.0
#define VI_WGS84_F (1.0 / 298.257223563)
#define VI_RE 6371200.0  // sphharmlag.py:9
#define VI_PI 3.14159265358979323846

VI_HD void vi_geodetic2ecef(double lat_deg, double lon_deg, double alt, double* x, double* y, double* z) {
  const double b = VI_WGS84_A * (1.0 - VI_WGS84_F);
  double lat = lat_deg * (VI_PI / 180.0);
  double lon = lon_deg * (VI_PI / 180.0);
  double sl = sin(lat), cl = cos(lat);
  double n = (VI_WGS84_A * VI_WGS84_A) / hypot(VI_WGS84_A * cl, b * sl);
  *x = (n + alt) * cl * cos(lon);
  *y = (n + alt) * cl * sin(lon);
  double ba = b / VI_WGS84_A;
  *z = (n * (ba * ba) + alt) * sl;
}

// (z, theta, phi) of sphharmlag.py:353-359 — Rodrigues rotation by +theta0 about k.
VI_HD void vi_shl_coords(const vi_shl_params& P, double lat, double lon, double alt,
                         double* zz, double* theta, double* phi) {
  double x, y, z;
  vi_geodetic2ecef(lat, lon, alt, &x, &y, &z);
  double omc = 1.0 - P.ct0;
  double kd = P.kx * x + P.ky * y;   // k.R (k_z = 0)
  double cx = P.ky * z;              // (k x R)_x = ky*Rz - 0*Ry
  double cy = -(P.kx * z);           // (k x R)_y = 0*Rx - kx*Rz
  double cz = P.kx * y - P.ky * x;   // (k x R)_z
  double rx = x * P.ct0 + cx * P.st0 + P.kx * kd * omc;
  double ry = y * P.ct0 + cy * P.st0 + P.ky * kd * omc;
  double rz = z * P.ct0 + cz * P.st0;
  double r = sqrt(rx * rx + ry * ry + rz * rz);
  *theta = acos(rz / r);
  *phi = atan2(ry, rx);
  *zz = 100.0 * (r / VI_RE - 1.0);
}

// scipy eval_laguerre for integer order: forward recurrence on the increments.
// out[k] = L_k(x), k = 0..maxk-1.
VI_HD void vi_laguerre_all(int maxk, double x, double* out) {
  out[0] = 1.0;
  if (maxk > 1) out[1] = 1.0 - x;   // (-x) + 0 + 1
  double d = -x;                    // -x/(alpha+1)
  double p = d + 1.0;
  for (int kk = 0; kk + 2 < maxk; ++kk) {
    double k = kk + 1.0;
    d = -x / (k + 1.0) * p + (k / (k + 1.0)) * d;
    p = d + p;
    out[kk + 2] = p;
  }
}

// Zhang & Jin LPMV0: P_v^m(x) for real v >= 0, integer m >= 0, -0.35 <= x <= 1.
VI_HD double vi_lpmv0(double v, int m, double x) {
  int nv = (int)v;
  double v0 = v - nv;
  double c0 = 1.0;
  if (m != 0) {
    double rg = v * (v + m);
    for (int j = 1; j < m; ++j) rg = rg * (v * v - (double)(j * j));
    double xq = sqrt(1.0 - x * x);
    double r0 = 1.0;
    for (int j = 1; j <= m; ++j) r0 = 0.5 * r0 * xq / j;
    c0 = r0 * rg;
  }
  if (v0 == 0.0) {
    double pmv = 1.0, r = 1.0;
    for (int k = 1; k <= nv - m; ++k) {
      r = 0.5 * r * (-nv + m + k - 1.0) * (nv + m + k) / (double)(k * (k + m)) * (1.0 + x);
      pmv += r;
    }
    return ((nv & 1) ? -1.0 : 1.0) * c0 * pmv;
  }
  double pmv = 1.0, r = 1.0;
  for (int k = 1; k <= 100; ++k) {
    r = 0.5 * r * (-v + m + k - 1.0) * (v + m + k) / (double)(k * (m + k)) * (1.0 - x);
    pmv += r;
    if (k > 12 && fabs(r / pmv) < 1e-14) break;
  }
  return ((m & 1) ? -1.0 : 1.0) * c0 * pmv;
}

// Zhang & Jin LPMV for order |m| (the negative-order reflection is applied by
// the caller, which holds the Gamma tables).  Upward degree recurrence (AMS
// 8.5.3) from the two series seeds when int(v) > max(2, |m|).
VI_HD double vi_lpmv_pos(double v, int mx, double x) {
  int nv = (int)v;
  double v0 = v - nv;
  if (nv > 2 && nv > mx) {
    double p0 = vi_lpmv0(v0 + mx, mx, x);
    double p1 = vi_lpmv0(v0 + mx + 1, mx, x);
    double pmv = p1;
    for (int j = mx + 2; j <= nv; ++j) {
      pmv = ((2.0 * (v0 + j) - 1.0) * x * p1 - (v0 + j - 1.0 + mx) * p0) / (v0 + j - mx);
      p0 = p1;
      p1 = pmv;
    }
    return pmv;
  }
  return vi_lpmv0(v, mx, x);
}

// One row of the sphharmlag design matrix (sphharmlag.py:138-141):
//   A[n] = exp(-z/2) * L_k(z) * (K_{v|m|} * cos|sin(|m| phi)) * P_v^m(cos theta)
// n = k*maxl^2 + l*(l+1) + m.  `emit(n, value)` receives every basis value.
// [l0, l1): the degrees to emit (default all) -- kernels that give every (point, degree) pair its own thread call it
// with one degree; the values do not depend on the split.
template <class Emit>
VI_HD void vi_shl_row(const vi_shl_params& P, double lat, double lon, double alt, Emit emit, int l0 = 0, int l1 = -1) {
  double z, theta, phi;
  vi_shl_coords(P, lat, lon, alt, &z, &theta, &phi);
  double lag[VI_MAXK_MAX];
  vi_laguerre_all(P.maxk, z, lag);
  double ez = exp(-0.5 * z);
  double x = cos(theta);
  const int L2 = P.maxl * P.maxl;
  if (l1 < 0 || l1 > P.maxl) l1 = P.maxl;
  for (int l = l0; l < l1; ++l) {
    double v = P.nu[l];
    for (int am = 0; am <= l; ++am) {
      double ppos = vi_lpmv_pos(v, am, x);
      double cs = cos(am * phi);
      double kv = P.kvm[l][am];
      double azc = kv * cs;
      int r_pos = l * (l + 1) + am;
      for (int k = 0; k < P.maxk; ++k) emit(k * L2 + r_pos, ez * lag[k] * azc * ppos);
      if (am > 0) {
        // negative order: lpmv(-am, v, x) = P_v^{am} * G1/G2 * (-1)^am when |P| < 1e300
        double pneg = ppos;
        if (fabs(ppos) < 1.0e300) pneg = ppos * P.g1[l][am] / P.g2[l][am] * ((am & 1) ? -1.0 : 1.0);
        double azs = kv * sin(am * phi);
        int r_neg = l * (l + 1) - am;
        for (int k = 0; k < P.maxk; ++k) emit(k * L2 + r_neg, ez * lag[k] * azs * pneg);
      }
    }
  }
}

// scipy eval_genlaguerre(n, 1, x) for n = -1 .. maxk-2 (eval_genlaguerre_l: the eval_laguerre recurrence with
// alpha = 1, times binom(n + 1, n) = n + 1).  out[k] = L^{(1)}_{k-1}(x), out[0] = 0 (scipy returns 0 for n < 0).
VI_HD void vi_genlaguerre1_all(int maxk, double x, double* out) {
  out[0] = 0.0;
  if (maxk > 1) out[1] = 1.0;            // n = 0
  if (maxk > 2) out[2] = -x + 1.0 + 1.0; // n = 1: -x + alpha + 1
  double d = -x / 2.0;                   // -x / (alpha + 1)
  double p = d + 1.0;
  for (int kk = 0; kk + 3 < maxk; ++kk) {
    double k = kk + 1.0;
    d = -x / (k + 2.0) * p + (k / (k + 2.0)) * d;
    p = d + p;
    out[kk + 3] = (kk + 3.0) * p;        // n = kk + 2: binom(n + 1, n) = n + 1
  }
}

// One row of the GRADIENT of the sphharmlag basis (sphharmlag.py:148-184), components along z-hat, theta-hat,
// phi-hat:
//   gz = -0.5 e (L_k + 2 L^{(1)}_{k-1}) P_v^m A_vm 100 / RE
//   gt = e L_k (-(v+1) x P_v^m + (v-m+1) P_{v+1}^m) A_vm / (y (z/100 + 1) RE)
//   gp = e L_k P_v^m A'_vm / (y (z/100 + 1) RE)
// with x = cos theta, y = sin theta, e = exp(-z/2), m signed.  `emit(n, gz, gt, gp)` receives every basis index.
// P_{v+1}^{-|m|} uses Gamma(v-|m|+2)/Gamma(v+|m|+2) = g1 (v-|m|+1) / (g2 (v+|m|+1)).
template <class Emit>
VI_HD void vi_shl_grad_row(const vi_shl_params& P, double lat, double lon, double alt, Emit emit) {
  double z, theta, phi;
  vi_shl_coords(P, lat, lon, alt, &z, &theta, &phi);
  double lag[VI_MAXK_MAX], lag1[VI_MAXK_MAX];
  vi_laguerre_all(P.maxk, z, lag);
  vi_genlaguerre1_all(P.maxk, z, lag1);
  const double e = exp(-0.5 * z);
  const double x = cos(theta), y = sin(theta);
  const double den = y * (z / 100.0 + 1.0) * VI_RE;
  const int L2 = P.maxl * P.maxl;
  for (int l = 0; l < P.maxl; ++l) {
    const double v = P.nu[l];
    for (int am = 0; am <= l; ++am) {
      const double ppos = vi_lpmv_pos(v, am, x);
      const double ppos1 = vi_lpmv_pos(v + 1.0, am, x);
      const double kv = P.kvm[l][am];
      const double cs = cos(am * phi), sn = sin(am * phi);
      {
        // m = +am: A = K cos(m phi), A' = -m K sin(m phi)
        const double a = kv * cs, da = -1.0 * am * kv * sn;
        const double tt = -(v + 1.0) * x * ppos + (v - am + 1.0) * ppos1;
        const int r = l * (l + 1) + am;
        for (int k = 0; k < P.maxk; ++k)
          emit(k * L2 + r, -0.5 * e * (lag[k] + 2.0 * lag1[k]) * ppos * a * 100.0 / VI_RE,
               e * lag[k] * tt * a / den, e * lag[k] * ppos * da / den);
      }
      if (am > 0) {
        // m = -am: A = K sin(|m| phi), A' = |m| K cos(|m| phi); negative-order Legendre by reflection
        const double sg = (am & 1) ? -1.0 : 1.0;
        double pneg = ppos, pneg1 = ppos1;
        if (fabs(ppos) < 1.0e300) pneg = ppos * P.g1[l][am] / P.g2[l][am] * sg;
        if (fabs(ppos1) < 1.0e300) pneg1 = ppos1 * (P.g1[l][am] * (v - am + 1.0)) / (P.g2[l][am] * (v + am + 1.0)) * sg;
        const double a = kv * sn, da = am * kv * cs;
        const double tt = -(v + 1.0) * x * pneg + (v + am + 1.0) * pneg1;
        const int r = l * (l + 1) - am;
        for (int k = 0; k < P.maxk; ++k)
          emit(k * L2 + r, -0.5 * e * (lag[k] + 2.0 * lag1[k]) * pneg * a * 100.0 / VI_RE,
               e * lag[k] * tt * a / den, e * lag[k] * pneg * da / den);
      }
    }
  }
}

// Gaussian RBF value for one (point, centre): radbasfun.py:106-107 takes
// r = ||R - c||_2 (np.linalg.norm) and then squares it again.
VI_HD double vi_rbf_value(double px, double py, double pz, double cx, double cy, double cz, double eps) {
  double dx = px - cx, dy = py - cy, dz = pz - cz;
  double r = sqrt(dx * dx + dy * dy + dz * dz);
  return exp(-(r * r) / (eps * eps));
}
