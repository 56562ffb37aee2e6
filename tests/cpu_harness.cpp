// TEST-ONLY: compiles the device-agnostic (VI_HD) scalar routines of
// volumetricinterp_b200/csrc for the CPU so that their arithmetic can be checked
// against scipy / LAPACK in the GPU-less build container.  This shared object is
// never loaded by the product package; the product path fails loudly without the
// CUDA library (volumetricinterp_b200/_native.py).
#include <cstdint>
#include <cstring>
#include <vector>
#define VI_EMU 1
#include "../volumetricinterp_b200/csrc/vi_math.h"
#include "../volumetricinterp_b200/csrc/vi_tql.h"
#include "../volumetricinterp_b200/csrc/vi_brent.h"
#include "../volumetricinterp_b200/csrc/vi_tridiag.h"
#include "../volumetricinterp_b200/csrc/vi_tridiag_packed.h"
#include "../volumetricinterp_b200/csrc/vi_nm.h"
#include "../volumetricinterp_b200/csrc/vi_ne_split.h"
#include "../volumetricinterp_b200/csrc/vi_band.h"
#include "../volumetricinterp_b200/csrc/vi_chase.h"
#include "../volumetricinterp_b200/csrc/vi_wave.h"

extern "C" {

void h_shl_rows(const vi_shl_params* P, const double* lat, const double* lon, const double* alt,
                int64_t npts, double* A) {
  const int N = P->maxk * P->maxl * P->maxl;
  for (int64_t p = 0; p < npts; ++p) {
    double* row = A + p * N;
    vi_shl_row(*P, lat[p], lon[p], alt[p], [&](int n, double v) { row[n] = v; });
  }
}

void h_shl_grad_rows(const vi_shl_params* P, const double* lat, const double* lon, const double* alt,
                     int64_t npts, double* out) {
  const int N = P->maxk * P->maxl * P->maxl;
  for (int64_t p = 0; p < npts; ++p) {
    double* o = out + p * 3 * (int64_t)N;
    vi_shl_grad_row(*P, lat[p], lon[p], alt[p], [&](int n, double gz, double gt, double gp) {
      o[n] = gz; o[N + n] = gt; o[2 * N + n] = gp;
    });
  }
}

void h_shl_coords(const vi_shl_params* P, const double* lat, const double* lon, const double* alt,
                  int64_t npts, double* z, double* th, double* ph) {
  for (int64_t p = 0; p < npts; ++p) vi_shl_coords(*P, lat[p], lon[p], alt[p], z + p, th + p, ph + p);
}

double h_lpmv_pos(double v, int m, double x) { return vi_lpmv_pos(v, m, x); }

void h_rbf_rows(const double* lat, const double* lon, const double* alt, int64_t npts,
                const double* centers, int N, double eps, double* A) {
  for (int64_t p = 0; p < npts; ++p) {
    double x, y, z;
    vi_geodetic2ecef(lat[p], lon[p], alt[p], &x, &y, &z);
    for (int n = 0; n < N; ++n)
      A[p * N + n] = vi_rbf_value(x, y, z, centers[3 * n], centers[3 * n + 1], centers[3 * n + 2], eps);
  }
}

// Truncated pseudo-inverse solve of a symmetric tridiagonal system:
// w = Z L^+ Z^T g0 ; lam <- eigenvalues ; returns status, *rank, *nrot.
int h_tql_solve(int n, const double* d0, const double* e0, const double* g0, double rcond,
                double* w, double* lam, int* rank, int* nrot) {
  std::vector<double> d(d0, d0 + n), e(n, 0.0), g(g0, g0 + n);
  for (int i = 0; i + 1 < n; ++i) e[i] = e0[i];
  int cap = 4 * n * n + 64;
  std::vector<double> tc(cap), ts(cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tc.data(), 1}, {ts.data(), 1}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  int st = vi_tql(n, {d.data(), 1}, {e.data(), 1}, {g.data(), 1}, tape, &nr);
  std::memcpy(lam, d.data(), n * sizeof(double));
  *rank = vi_spectral_divide(n, {d.data(), 1}, {g.data(), 1}, rcond);
  vi_tape_apply_z({g.data(), 1}, tape, nr);
  std::memcpy(w, g.data(), n * sizeof(double));
  *nrot = nr;
  return st;
}

// Same solve through the GPU hot-path variant: eigenvalues + tape (vi_tql_values), Z^T g by forward
// replay, divide, Z u by backward replay.
int h_tql_values_solve(int n, const double* d0, const double* e0, const double* g0, double rcond,
                       double* w, double* lam, int* rank, int* nrot) {
  std::vector<double> d(d0, d0 + n), e(n, 0.0), g(g0, g0 + n);
  for (int i = 0; i + 1 < n; ++i) e[i] = e0[i];
  int cap = 2 * n * n + 64;
  std::vector<double> tcs(2 * (size_t)cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tcs.data(), 2}, {tcs.data() + 1, 2}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  int st = vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr);
  std::memcpy(lam, d.data(), n * sizeof(double));
  vi_tape_apply_zt({g.data(), 1}, tape, nr);
  *rank = vi_spectral_divide(n, {d.data(), 1}, {g.data(), 1}, rcond);
  vi_tape_apply_z({g.data(), 1}, tape, nr);
  std::memcpy(w, g.data(), n * sizeof(double));
  *nrot = nr;
  return st;
}

// nested-loop QL (vi_tql_values) vs flattened state machine (vi_tql_values_flat): eigenvalues and tape
// must be bit-identical.  Returns 0 if identical, a positive code otherwise.
int h_tql_flat_identical(int n, const double* d0, const double* e0, int* nrot_out) {
  std::vector<double> d1(d0, d0 + n), e1(n, 0.0), d2(d0, d0 + n), e2(n, 0.0);
  for (int i = 0; i + 1 < n; ++i) { e1[i] = e0[i]; e2[i] = e0[i]; }
  int cap = 2 * n * n + 64;
  std::vector<double> t1(2 * (size_t)cap, 0.0), t2(2 * (size_t)cap, 0.0);
  std::vector<int32_t> i1(cap, 0), i2(cap, 0);
  vi_tape ta{{t1.data(), 2}, {t1.data() + 1, 2}, {i1.data(), 1}, cap};
  vi_tape tb{{t2.data(), 2}, {t2.data() + 1, 2}, {i2.data(), 1}, cap};
  int32_t n1 = 0, n2 = 0;
  int s1 = vi_tql_values(n, {d1.data(), 1}, {e1.data(), 1}, ta, &n1);
  int s2 = vi_tql_values_flat(n, {d2.data(), 1}, {e2.data(), 1}, tb, &n2, true);
  *nrot_out = n2;
  if (s1 != s2) return 1;
  if (n1 != n2) return 2;
  if (std::memcmp(d1.data(), d2.data(), n * sizeof(double)) != 0) return 3;
  if (std::memcmp(t1.data(), t2.data(), 2 * (size_t)n1 * sizeof(double)) != 0) return 4;
  if (std::memcmp(i1.data(), i2.data(), (size_t)n1 * sizeof(int32_t)) != 0) return 5;
  return 0;
}

typedef double (*h_fn)(double);
// brentq driven through the state machine; xs receives every abscissa evaluated.
int h_brentq(h_fn f, double xa, double xb, double* root, double* xs, int* nfev) {
  vi_brent b;
  int n = 0;
  double fa = f(xa), fb = f(xb);
  xs[n++] = xa; xs[n++] = xb;
  vi_brent_init(b, xa, fa, xb, fb);
  while (!vi_brent_propose(b)) {
    xs[n++] = b.xcur;
    vi_brent_feed(b, f(b.xcur));
  }
  *root = b.root;
  *nfev = n;
  return b.done;
}

// Nelder-Mead state machine driven by a callback; xs receives every abscissa evaluated.
int h_nm_minimize(h_fn f, double x0, double* xmin, double* xs, int* nfev, int* nit) {
  vi_nm s;
  vi_nm_init(s, x0);
  double x;
  int n = 0;
  while (vi_nm_next(s, &x)) {
    xs[n++] = x;
    vi_nm_feed(s, f(x));
  }
  *xmin = s.x0;
  *nfev = s.fcalls;
  *nit = s.iters;
  return s.success;
}

void h_chi2_bracket(const double* table, int npts, int* status, int* k_lo, double* nu) {
  vi_bracket br = vi_chi2_bracket(table, 1, npts);
  *status = br.status; *k_lo = br.k_lo; *nu = br.nu;
}

// Lazy table phase of vi_fit_batched (k_table_plan / k_table_advance in csrc/fit.cu): entries are revealed `step`
// at a time, entries past `lim` replicate entry lim-1; returns how many distinct entries had to be evaluated
// and the bracket found on the partially filled table (unevaluated entries are poisoned with NaN).
int h_chi2_bracket_lazy(const double* full, int npts, int step, int lim, int* status, int* k_lo, double* nu) {
  double tab[VI_NALPHA];
  for (int k = 0; k < VI_NALPHA; ++k) tab[k] = NAN;
  int kdone = 0;
  bool walking = true;
  while (walking) {
    int c = lim - kdone; if (c > step) c = step;
    if (c <= 0) break;
    for (int k = kdone; k < kdone + c; ++k) tab[k] = full[k];
    kdone += c;
    int avail = kdone;
    if (kdone >= lim) { for (int k = lim; k < VI_NALPHA; ++k) tab[k] = tab[lim - 1]; avail = VI_NALPHA; }
    walking = vi_chi2_walk_needs_more(tab, 1, npts, avail);
  }
  vi_bracket br = vi_chi2_bracket(tab, 1, npts);
  *status = br.status; *k_lo = br.k_lo; *nu = br.nu;
  return kdone;
}

// The complete per-system pipeline of kernels k_tridiag + k_tql (csrc/fit.cu) executed on the CPU
// with the phase bodies of vi_tridiag.h run for tid = 0..nt-1 in turn:
//   X = 2^-ex (sym(G) + sum lam_r Reg_r) -> T = Q^T X Q -> QL with tape -> truncated solve -> C = Q c~.
// Returns the QL status; *bad = 1 if a non-finite entry was met.
int h_system_solve(int n, const double* G, const double* y, const double* regs, const double* lam, int nreg,
                   int nt, double rcond, double* C, int* rank, double* dd, double* ee, int* bad) {
  const int ld = vi_tri_ld(n);
  std::vector<double> X((size_t)n * ld), aux(vi_tri_aux_doubles(n, nt) + 16), V((size_t)n * n, 0.0);
  vi_tri_ws S;
  double* a = aux.data();
  S.X = X.data(); S.ld = ld;
  vi_tri_carve(S, a, n, nt);
  vi_tri_load(S, n, G, y, regs, lam, nreg, 0, nt);
  *bad = S.sc[1] != 0.0;
  if (*bad) return 0;
  vi_tri_reduce(S, n, V.data(), 0, nt);
  for (int i = 0; i < n; ++i) { dd[i] = S.d[i]; ee[i] = S.e[i]; }
  std::vector<double> d(S.d, S.d + n), e(S.e, S.e + n), g(S.yv, S.yv + n), tau(S.tau, S.tau + n);
  int cap = 2 * n * n + 64;
  std::vector<double> tcs(2 * (size_t)cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tcs.data(), 2}, {tcs.data() + 1, 2}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  int st = vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr);
  if (st != 0) return st;
  vi_tape_apply_zt({g.data(), 1}, tape, nr);
  *rank = vi_spectral_divide(n, {d.data(), 1}, {g.data(), 1}, rcond);
  vi_tape_apply_z({g.data(), 1}, tape, nr);
  for (int i = 0; i < n; ++i) g[i] *= S.sc[0];
  vi_tri_backtransform(n, V.data(), tau.data(), 1, g.data(), 1);
  std::memcpy(C, g.data(), n * sizeof(double));
  return 0;
}

// Same pipeline with the packed-triangle tridiagonalisation (vi_tridiag_packed.h, kernel k_tridiag_packed):
// phases run thread by thread, warp reductions restated in the device's order.
int h_system_solve_packed(int n, const double* G, const double* y, const double* regs, const double* lam, int nreg,
                          double rcond, double* C, int* rank, double* dd, double* ee, int* bad) {
  const int nt = vi_trp_threads(n);
  std::vector<double> mem(vi_trp_doubles(n) + 16), V((size_t)n * n, 0.0);
  vi_trp_ws W;
  double* base = mem.data();
  if (reinterpret_cast<uintptr_t>(base) & 15) base += 1;
  vi_trp_carve(W, base, n);
  vi_trp_load(W, G, y, regs, lam, nreg, 0, nt);
  *bad = W.sc[1] != 0.0;
  if (*bad) return 0;
  vi_trp_reduce(W, V.data(), 0, nt);
  for (int i = 0; i < n; ++i) { dd[i] = W.d[i]; ee[i] = W.e[i]; }
  std::vector<double> d(W.d, W.d + n), e(W.e, W.e + n), g(W.yv, W.yv + n), tau(W.tau, W.tau + n);
  int cap = 2 * n * n + 64;
  std::vector<double> tcs(2 * (size_t)cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tcs.data(), 2}, {tcs.data() + 1, 2}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  int st = vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr);
  if (st != 0) return st;
  vi_tape_apply_zt({g.data(), 1}, tape, nr);
  *rank = vi_spectral_divide(n, {d.data(), 1}, {g.data(), 1}, rcond);
  vi_tape_apply_z({g.data(), 1}, tape, nr);
  for (int i = 0; i < n; ++i) g[i] *= W.sc[0];
  vi_tri_backtransform(n, V.data(), tau.data(), 1, g.data(), 1);
  std::memcpy(C, g.data(), n * sizeof(double));
  return 0;
}

// Work split of k_ne_dmma3: fills unit (mi | ni << 8 | flags), wbeg / wend per warp; returns 1 if it fits.
int h_ne3_split(int N, int DT, int* nunits, int* uinfo, int* wbeg, int* wend) {
  NeSplit sp;
  const int mt = (N + 15) / 16;
  if (!ne3_split(N, mt, DT, sp)) return 0;
  *nunits = sp.nunits;
  for (int i = 0; i < sp.nunits; ++i) uinfo[i] = sp.uinfo[i];
  for (int w = 0; w < kW3; ++w) { wbeg[w] = sp.wbeg[w]; wend[w] = sp.wend[w]; }
  return 1;
}
int h_ne3_warps() { return kW3; }

// Packed-triangle layout of k_tridiag_packed (vi_tridiag_packed.h)
int h_trp_geometry(int n, int* npad, int* noct, int* nwarp, int* xdoubles, int* ntiles) {
  *npad = vi_trp_npad(n); *noct = vi_trp_noct(n); *nwarp = vi_trp_nwarp(n);
  *xdoubles = vi_trp_xdoubles(n); *ntiles = vi_trp_ntiles(n);
  return vi_trp_doubles(n);
}
int h_trp_idx(int n, int i, int c) { return vi_trp_idx(vi_trp_npad(n), i, c); }
int h_trp_tile_slot(int n, int ip, int q) { return vi_trp_tileoff(vi_trp_npad(n), q) + ip - 4 * q; }
int h_trp_octet_of(int n, int warp, int slot, int a) {
  vi_trp_ws W;
  W.n = n; W.npad = vi_trp_npad(n); W.noct = vi_trp_noct(n); W.nwarp = vi_trp_nwarp(n);
  return vi_trp_octet_of(W, warp, slot, a);
}

static int g_split_p1 = 0, g_split_nw = 0, g_split_qt = 4, g_big = 0;
void h_bnd_set_big(int on) { g_big = on; }
void h_bnd_set_split(int p1, int nw2, int qt) { g_split_p1 = p1; g_split_nw = nw2; g_split_qt = qt; }

// Two-stage pipeline of kernels k_band + k_chase + QL + back-transformation (csrc/vi_band.h, vi_chase.h), the device
// code itself executed by the CUDA-on-CPU model of tests/cuda_emu.h (one fiber per CUDA thread):
//   X -> band (one CTA) -> tridiagonal (one warp) -> QL with tape -> truncated solve -> C = Q1 Q2 c~.
// bandout (optional): 9 * npad doubles, the band stage 1 produced.
int h_system_solve_two_stage(int n, const double* G, const double* y, const double* regs, const double* lam, int nreg,
                             double rcond, double* C, int* rank, double* dd, double* ee, int* bad, double* bandout) {
  const int nt = vi_bnd_threads(n), np = vi_bnd_npad(n);
  std::vector<double> smem(vi_bnd_doubles(n) + 2), Vg(vi_bnd_vdoubles(n) + 64, 0.0), band(vi_bnd_band_doubles(n), 0.0);
  std::vector<double> refl(vi_chs_rdoubles(n) + 8, 0.0);
  double* base = smem.data();
  if (reinterpret_cast<uintptr_t>(base) & 15) base += 1;
  double scl = 1.0, badf = 0.0;
  const int p1 = (g_split_p1 > 0 && g_split_p1 < vi_bnd_nbk(n) - 1) ? g_split_p1 : 0;      // split reduction (k_band + k_band_tail)
  std::vector<double> Xt(p1 ? vi_bnd_trailing_doubles(n, p1) + 2 : 2, 0.0);
  std::vector<double> Xbig(g_big ? vi_bnd_nblk(n) * 64 + 2 : 2, 0.0);
  if (g_big) {          // k_band_big: blocks in "global" memory, panel QR in shared memory, no look-ahead
    const int nwb = 6;
    std::vector<double> smemb(vi_bnd_doubles_big(n, nwb) + 2);
    double* bb = smemb.data();
    if (reinterpret_cast<uintptr_t>(bb) & 15) bb += 1;
    emu::run_cta(0, 32 * nwb, [&]() {
      vi_bnd_ws S;
      vi_bnd_carve_big(S, bb, Xbig.data(), n, nwb);
      vi_bnd_load(S, G, y, regs, lam, nreg, nullptr, 0.0, 0.0);
      const bool isbad = S.sc[1] != 0.0;
      if (!isbad) { vi_bnd_reduce_big(S, Vg.data()); vi_bnd_store_band(S, band.data()); }
      if (vi_tid() == 0) { scl = S.sc[0]; badf = S.sc[1]; }
    });
  } else
  emu::run_cta(0, nt, [&]() {
    vi_bnd_ws S;
    vi_bnd_carve(S, base, n);
    vi_bnd_load(S, G, y, regs, lam, nreg, nullptr, 0.0, 0.0);
    const bool isbad = S.sc[1] != 0.0;
    if (!isbad) {
      if (p1) {
        vi_bnd_reduce(S, Vg.data(), p1);
        vi_bnd_store_band(S, band.data(), band.data() + 9 * np, 8 * p1);
        vi_bnd_store_trailing(S, p1, Xt.data());
      } else {
        vi_bnd_reduce(S, Vg.data());
        vi_bnd_store_band(S, band.data());
      }
    }
    if (vi_tid() == 0) { scl = S.sc[0]; badf = S.sc[1]; }
  });
  if (p1 && !g_big && badf == 0.0) {
    const int n2 = n - 8 * p1, nw2 = g_split_nw > 0 ? g_split_nw : vi_bnd_nwarp(n2);
    std::vector<double> smem2(vi_bnd_doubles(n2, nw2) + 2);
    double* base2 = smem2.data();
    if (reinterpret_cast<uintptr_t>(base2) & 15) base2 += 1;
    emu::run_cta(0, 32 * nw2, [&]() {
      vi_bnd_ws S;
      vi_bnd_carve(S, base2, n2, nw2);
      vi_bnd_load_trailing(S, Xt.data());
      double* Vg2 = Vg.data() + vi_bnd_voff(np, p1);
      if (g_split_qt == 3 && S.npad - 8 <= 96) vi_bnd_reduce<3>(S, Vg2);
      else vi_bnd_reduce<4>(S, Vg2);
      vi_bnd_store_band(S, band.data() + 9 * 8 * p1, band.data() + 9 * np + 8 * p1, S.npad);
    });
  }
  *bad = badf != 0.0;
  if (*bad) return 0;
  if (bandout) std::memcpy(bandout, band.data(), 9 * (size_t)np * sizeof(double));
  std::vector<double> work(vi_chs_doubles(n) + 2, 0.0);
  double* Bw = work.data();
  double* gw = Bw + VI_CHS_LDB * np;
  emu::run_cta(0, 32, [&]() {
    vi_chs_load(Bw, gw, band.data(), n);
    vi_chs_reduce(Bw, gw, n, refl.data());
  });
  std::vector<double> d(n), e(n, 0.0), g(gw, gw + n);
  for (int j = 0; j < n; ++j) { d[j] = Bw[j * VI_CHS_LDB]; if (j + 1 < n) e[j] = Bw[j * VI_CHS_LDB + 1]; }
  for (int i = 0; i < n; ++i) { dd[i] = d[i]; ee[i] = e[i]; }
  int cap = 2 * n * n + 64;
  std::vector<double> tcs(2 * (size_t)cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tcs.data(), 2}, {tcs.data() + 1, 2}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  int st = vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr);
  if (st != 0) return st;
  vi_tape_apply_zt({g.data(), 1}, tape, nr);
  *rank = vi_spectral_divide(n, {d.data(), 1}, {g.data(), 1}, rcond);
  vi_tape_apply_z({g.data(), 1}, tape, nr);
  for (int i = 0; i < n; ++i) g[i] *= scl;
  std::vector<double> u(np + 8, 0.0);
  std::memcpy(u.data(), g.data(), n * sizeof(double));
  emu::run_cta(0, 32, [&]() {
    vi_chs_apply_q(u.data(), n, refl.data());
    vi_bnd_apply_q(u.data(), n, Vg.data());
  });
  std::memcpy(C, u.data(), n * sizeof(double));
  return 0;
}
// Wavefront tape replay (vi_wave.h) against the sequential replay (vi_tql.h): QL of the tridiagonal (d0, e0), then
// Z^T g and Z g both ways.  Returns 0 if the results are BIT-IDENTICAL, a positive code otherwise; *nsweeps = sweeps found.
int h_wave_replay_identical(int n, const double* d0, const double* e0, const double* g0, int* nsweeps, int* nrot_out) {
  std::vector<double> d(d0, d0 + n), e(n, 0.0);
  for (int i = 0; i + 1 < n; ++i) e[i] = e0[i];
  int cap = 2 * n * n + 64;
  std::vector<double> tcs(2 * (size_t)cap);
  std::vector<int32_t> ti(cap);
  vi_tape tape{{tcs.data(), 2}, {tcs.data() + 1, 2}, {ti.data(), 1}, cap};
  int32_t nr = 0;
  if (vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr) != 0) return 9;
  *nrot_out = nr;
  std::vector<double> a(g0, g0 + n), b(g0, g0 + n), wa(g0, g0 + n), wb(g0, g0 + n);
  vi_tape_apply_zt({a.data(), 1}, tape, nr);
  vi_tape_apply_z({b.data(), 1}, tape, nr);
  const int maxsw = vi_wav_maxsweeps(n);
  std::vector<int32_t> tab(maxsw + 2);
  int ns = 0, tnext = 0;
  emu::run_cta(0, 32, [&]() {
    int tn = 0;
    const int k = vi_wav_scan(ti.data(), 0, nr, tab.data(), maxsw, &tn);
    if (vi_tid() == 0) { ns = k; tnext = tn; }
    if (tn == nr) {
      vi_wav_pass<true>(wa.data(), tcs.data(), ti.data(), tab.data(), k);
      vi_wav_pass<false>(wb.data(), tcs.data(), ti.data(), tab.data(), k);
    }
  });
  *nsweeps = ns;
  if (tnext != nr) return 8;            // table overflow (the kernel then replays sequentially)
  if (std::memcmp(a.data(), wa.data(), n * sizeof(double)) != 0) return 1;
  if (std::memcmp(b.data(), wb.data(), n * sizeof(double)) != 0) return 2;
  return 0;
}

// The same with 32 / G systems replayed side by side by one warp (lane groups of G = 8 or 16): `nsys` tridiagonals of
// order n, concatenated.  Returns 0 if every system's Z^T g and Z g are bit-identical to the sequential replay.
int h_wave_replay_groups(int n, int G, const double* d0, const double* e0, const double* g0) {
  const int S = 32 / G;
  const int cap = 2 * n * n + 64, maxsw = vi_wav_maxsweeps(n);
  std::vector<std::vector<double>> tcs(S), a(S), b(S), wa(S), wb(S);
  std::vector<std::vector<int32_t>> ti(S), tab(S);
  std::vector<int32_t> nr(S, 0);
  std::vector<int> ns(S, 0);
  for (int q = 0; q < S; ++q) {
    std::vector<double> d(d0 + q * n, d0 + (q + 1) * n), e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = e0[q * n + i];
    tcs[q].assign(2 * (size_t)cap, 0.0); ti[q].assign(cap, 0); tab[q].assign(maxsw + 2, 0);
    vi_tape tape{{tcs[q].data(), 2}, {tcs[q].data() + 1, 2}, {ti[q].data(), 1}, cap};
    if (vi_tql_values(n, {d.data(), 1}, {e.data(), 1}, tape, &nr[q]) != 0) return 9;
    a[q].assign(g0 + q * n, g0 + (q + 1) * n); b[q] = a[q]; wa[q] = a[q]; wb[q] = a[q];
    vi_tape_apply_zt({a[q].data(), 1}, tape, nr[q]);
    vi_tape_apply_z({b[q].data(), 1}, tape, nr[q]);
  }
  int bad = 0;
  emu::run_cta(0, 32, [&]() {
    const int lane = vi_tid() & 31, grp = lane / G;
    int myns = 0;
    for (int q = 0; q < S; ++q) {              // the scans: whole warp, one system after the other
      int tn = 0;
      const int k = vi_wav_scan(ti[q].data(), 0, nr[q], tab[q].data(), maxsw, &tn);
      if (tn != nr[q]) { if (lane == 0) bad = 8; }
      if (q == grp) myns = k;
      if (lane == 0) ns[q] = k;
    }
    if (bad) return;
    if (G == 8) {
      vi_wav_pass<true, 8>(wa[grp].data(), tcs[grp].data(), ti[grp].data(), tab[grp].data(), myns);
      vi_wav_pass<false, 8>(wb[grp].data(), tcs[grp].data(), ti[grp].data(), tab[grp].data(), myns);
    } else {
      vi_wav_pass<true, 16>(wa[grp].data(), tcs[grp].data(), ti[grp].data(), tab[grp].data(), myns);
      vi_wav_pass<false, 16>(wb[grp].data(), tcs[grp].data(), ti[grp].data(), tab[grp].data(), myns);
    }
  });
  if (bad) return bad;
  for (int q = 0; q < S; ++q) {
    if (std::memcmp(a[q].data(), wa[q].data(), n * sizeof(double)) != 0) return 10 + q;
    if (std::memcmp(b[q].data(), wb[q].data(), n * sizeof(double)) != 0) return 20 + q;
  }
  return 0;
}

int h_wav_bytes(int n) { return vi_wav_bytes(n); }
int h_bnd_threads(int n) { return vi_bnd_threads(n); }
int h_bnd_smem_bytes(int n) { return vi_bnd_doubles(n) * 8; }
int h_chs_nrefl(int n) { return vi_chs_nrefl(n); }
int h_chs_off(int n, int s) { return vi_chs_off(n, s); }
int h_bnd_el(int r, int c) { return vi_bnd_el(r, c); }

int h_sizeof_shl_params() { return (int)sizeof(vi_shl_params); }
}
