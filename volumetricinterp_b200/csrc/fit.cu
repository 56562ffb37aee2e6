// K3 — per-record solve and regularisation-parameter search, batched over records.
//
// Replaces, for R records at once, the body of the reference's record loop after the normal
// equations (interpolate.py:555-569):
//   find_reg_param -> chi2 -> chi2objfunct -> eval_C      interpolate.py:97-147, 152-261, 432-469
//   NaN-record rule                                         interpolate.py:558-563
//   final eval_C and chi^2                                  interpolate.py:566, 569
//
// The reference re-solves (A^T W A + 10^alpha R) C = A^T W b with scipy.linalg.lstsq (LAPACK
// gelsd, rcond = eps) ~400-500 times per record, recomputing A^T W A each time.  Here the normal
// equations come from K2 once per record, and every trial is one "system":
//   k_tridiag  one CTA per system: X = sym(G) + sum lambda R in shared memory, Householder
//              tridiagonalisation, g = Q^T y                                  (vi_tridiag.h)
//   k_tql      one THREAD per system: implicit QL with a rotation tape, truncated spectral
//              solve (|eig| > eps*max|eig| == gelsd's rcond rule), back-transform with the
//              stored reflectors                                              (vi_tql.h)
//   k_chi2     chi^2 = sum_j W_j ((A C)_j - b_j)^2 exactly as chi2objfunct does (residual form),
//              16 systems per CTA so the design matrix is streamed once per 16 systems
// chi2objfunct(alpha) - nu depends on the scale factor only through nu, so the decade walk of
// interpolate.py:180-207 runs on a table chi2(10^-k), k = 0..101, evaluated ONCE per record
// (all its systems are independent -> one big batch); Brent's iterations are then advanced for
// all records in lock step (vi_brent.h state machine), one batch of systems per iteration.
#include "common.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <vector>
#include "vi_brent.h"
#include "vi_nm.h"
#include "vi_tql.h"
#include "vi_tridiag.h"
#include "vi_tridiag_packed.h"
#include "vi_band.h"
#include "vi_chase.h"
#include "vi_wave.h"

namespace {

constexpr int kSkip = -1;          // system slot not in use this round
constexpr int kChiSB = 16;         // systems per CTA in k_chi2
constexpr int kChiThreads = 256;

// ------------------------------------------------------------------------------------------
// workspace carving (identical arithmetic in vi_fit_workspace_bytes and the entry points)
// ------------------------------------------------------------------------------------------
struct Bump {
  char* base; int64_t off; int64_t cap;
  template <class T> T* take(int64_t count) {
    off = vi_align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * (int64_t)sizeof(T);
    return p;
  }
};

struct SysBuf {
  int64_t cap;         // systems held at once (multiple of 32)
  int n, nreg, tapecap, nt, use_gx, ld, nchunk;
  int two_stage;       // tridiagonalisation by band reduction + bulge chasing (vi_band.h, vi_chase.h)
  int split_p1;        // > 0: stage 1 in two kernels, the panels [0, split_p1) in k_band, the rest in k_band_tail
  int64_t vstride;     // doubles of reflector storage per system
  size_t smem;
  double *V, *d, *e, *g, *tau, *scl, *tcs, *lam, *Csys, *chi2, *chi2p, *Xg, *band;
  int32_t *tix, *st, *rec, *rank, *nrot, *unit, *kidx, *gate;
  int32_t* wtab;       // sweep tables of k_replay_wave4 (4n + 34 per system)
  unsigned long long* rot_total;   // rotations of all eigen-solves since the call started (profiling; may be null)
};

// leave-one-gate-out systems (GCV): system s is built from G - w a a^T, y - w b a with a = A[gate[s]]
struct Downdate { const double* A; const double* Wm; const double* bm; int P; };

constexpr int kChiGates = 1024;    // gates per CTA in k_chi2 (partial sums combined in fixed order)

// two-stage reduction: orders whose panel fits the QR warp's registers and whose blocks fit one CTA's shared memory
// two-stage tridiagonalisation with X in shared memory (n <= 168) ...
bool two_stage_smem(int n) {
  return n <= VI_BND_NMAX && (size_t)vi_bnd_doubles(n) * sizeof(double) <= 227 * 1024;
}
// ... or with X in global memory (k_band_big: the high-order model of BASELINE configs[2], N = 500)
constexpr int kBigWarps = 6;
constexpr int kBigNmax = 1024;
bool two_stage_big(int n) {
  return !two_stage_smem(n) && n <= kBigNmax && (size_t)vi_bnd_doubles_big(n, kBigWarps) * sizeof(double) <= 227 * 1024 &&
         (size_t)vi_chs_doubles(n) * sizeof(double) <= 227 * 1024;
}
bool two_stage_ok(int n) {
  static const bool off = getenv("VI_ONE_STAGE") != nullptr;
  return !off && (two_stage_smem(n) || two_stage_big(n));
}

// Split of stage 1 (k_band + k_band_tail).  The reduction is a chain of panel steps whose latency barely depends on
// the size of the trailing matrix, while its shared-memory footprint (two CTAs per SM at n = 144) is set by the FIRST
// panel.  From the panel on at which the trailing matrix fits four CTAs per SM (order 88: 50 KB with 4 warps), a
// second kernel with that footprint takes over: twice as many systems in flight for the remaining panels.
constexpr int kTailWarps = 4;
int band_split(int n) {
  static const int env = getenv("VI_BAND_SPLIT") ? atoi(getenv("VI_BAND_SPLIT")) : -1;       // 0: off, > 0: that panel
  if (!two_stage_ok(n) || !two_stage_smem(n) || env == 0) return 0;
  const int nbk = vi_bnd_nbk(n);
  if (vi_bnd_nwarp(n) != VI_BND_NW) return 0;                // small orders: one kernel
  int p1 = 0;
  if (env > 0) p1 = env;
  else
    for (int p = 1; p < nbk - 2; ++p)
      if ((size_t)vi_bnd_doubles(n - 8 * p, kTailWarps) * sizeof(double) + 1024 <= (227 * 1024) / 4) { p1 = p; break; }
  if (p1 < 1 || p1 > nbk - 3) return 0;
  if (vi_bnd_npad(n - 8 * p1) - 8 > 96) return 0;            // the tail's panel QR holds 3 x 32 rows in registers
  return p1;
}

int tri_threads(int n) {
  const int npair = vi_tri_npair(n);
  int ng = 1024 / npair;
  if (ng > 8) ng = 8;
  if (ng < 1) ng = 1;
  int nt = (ng * npair + 31) / 32 * 32;
  if (nt > 1024) nt = 1024;
  if (nt < ((n + 31) & ~31)) nt = (n + 31) & ~31;     // one thread per row/column index is assumed by phases A, C
  return nt;
}

void sysbuf_carve(Bump& b, SysBuf& S, int64_t cap, int n, int nreg, int P) {
  S.cap = cap; S.n = n; S.nreg = nreg;
  S.tapecap = 2 * n * n + 64;   // generic spectra need ~1.4 n^2 rotations (measured), graded ones far fewer
  S.nt = tri_threads(n);
  S.ld = vi_tri_ld(n);
  S.nchunk = (P + kChiGates - 1) / kChiGates;
  size_t smem_x = (size_t)n * S.ld * sizeof(double);
  size_t smem_aux = (size_t)vi_tri_aux_doubles(n, S.nt) * sizeof(double);
  S.use_gx = (smem_x + smem_aux > 227 * 1024) ? 1 : 0;
  S.smem = S.use_gx ? smem_aux : smem_x + smem_aux;
  S.two_stage = two_stage_ok(n) ? 1 : 0;
  S.split_p1 = band_split(n);
  S.vstride = (int64_t)n * n;
  if (S.two_stage) {
    const int64_t need = (int64_t)vi_bnd_vdoubles(n) + vi_chs_rdoubles(n) + 8;
    if (need > S.vstride) S.vstride = need;
  }
  S.V = b.take<double>(cap * S.vstride);
  S.band = S.two_stage ? b.take<double>(cap * (int64_t)vi_bnd_band_doubles(n)) : nullptr;
  S.d = b.take<double>(cap * n);
  S.e = b.take<double>(cap * n);
  S.g = b.take<double>(cap * n);
  S.tau = b.take<double>(cap * n);
  S.scl = b.take<double>(cap);
  S.tcs = b.take<double>(cap * S.tapecap * 2);     // (c, s) pairs, contiguous per system
  S.tix = b.take<int32_t>(cap * S.tapecap);
  S.lam = b.take<double>(cap * (nreg > 0 ? nreg : 1));
  S.Csys = b.take<double>(cap * n);
  S.chi2 = b.take<double>(cap);
  S.chi2p = b.take<double>(cap * (S.nchunk > 0 ? S.nchunk : 1));
  S.st = b.take<int32_t>(cap);
  S.rec = b.take<int32_t>(cap);
  S.rank = b.take<int32_t>(cap);
  S.nrot = b.take<int32_t>(cap);
  S.unit = b.take<int32_t>(cap);
  S.kidx = b.take<int32_t>(cap);
  S.gate = b.take<int32_t>(cap);
  static const bool tab_smem = getenv("VI_WAVE_TAB_SMEM") != nullptr;      // A/B: sweep tables in shared memory
  S.wtab = tab_smem ? nullptr : b.take<int32_t>(cap * (int64_t)(vi_wav_maxsweeps(n) + 2));
  S.Xg = (S.use_gx || (S.two_stage && !two_stage_smem(n))) ? b.take<double>(cap * n * S.ld) : nullptr;
  S.rot_total = b.take<unsigned long long>(4);
}

constexpr int kCovChunk = 512;     // records whose eigenvector / H / T scratch is held at once

// side stream + events of the covariance host sink (per device; created on first use, kept for the process)
struct CovSink {
  cudaStream_t copy = nullptr;
  cudaEvent_t ready[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
  int init() {
    if (copy) return VI_OK;
    VI_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      VI_CUDA(cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming));
      VI_CUDA(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
    }
    return VI_OK;
  }
};
CovSink& cov_sink() {
  static CovSink sinks[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return sinks[(dev >= 0 && dev < 64) ? dev : 0];
}

struct UnitBuf {      // one search unit = (record, regulariser)
  double* table;      // U x VI_NALPHA
  vi_brent* br;       // U
  double* nu;         // U
  int32_t* status;    // U
  int32_t* active;    // U
  int32_t* tabbad;    // U
  int32_t* kstar;     // U
  int32_t* kdone;     // U   table entries chi2(10^-k), k < kdone, evaluated so far
  int32_t* walking;   // U   the decade walk of this unit still needs entries beyond kdone
  int64_t* off;       // U + 1: exclusive prefix sum of the number of table systems per unit in this pass
  vi_nm* nm;          // U   Nelder-Mead state (GCV)
  double* fsum;       // U   GCV objective being accumulated
  double* alpha;      // U   abscissa under evaluation
  int32_t* count;     // 1
  int32_t* klo;       // U   bracket decade found by the walk (table index of its lower end; -1: none)
};

void unit_carve(Bump& b, UnitBuf& Ub, int64_t U) {
  Ub.table = b.take<double>(U * VI_NALPHA);
  Ub.br = b.take<vi_brent>(U);
  Ub.nu = b.take<double>(U);
  Ub.status = b.take<int32_t>(U);
  Ub.active = b.take<int32_t>(U);
  Ub.tabbad = b.take<int32_t>(U);
  Ub.kstar = b.take<int32_t>(U);
  Ub.kdone = b.take<int32_t>(U);
  Ub.walking = b.take<int32_t>(U);
  Ub.off = b.take<int64_t>(U + 1);
  Ub.nm = b.take<vi_nm>(U);
  Ub.fsum = b.take<double>(U);
  Ub.alpha = b.take<double>(U);
  Ub.count = b.take<int32_t>(8);
  Ub.klo = b.take<int32_t>(U);
}

int64_t cov_scratch_bytes(int64_t R, int n) {
  int64_t cc = R < kCovChunk ? R : kCovChunk;
  // E, H, T and a two-chunk ring for the covariance when it is streamed to a pinned host buffer
  return (5 * cc * (int64_t)n * n + cc * n) * (int64_t)sizeof(double) + 8192;
}

int64_t gcv_scratch_bytes(int64_t R, int64_t P, int64_t U) { return R * P * 4 + R * 4 + U * 8 + 4096; }

int64_t per_system_bytes(int n, int nreg, int P) {
  Bump b{nullptr, 0, 0};
  SysBuf S;
  sysbuf_carve(b, S, 32, n, nreg, P);
  return (b.off + 32 * 256) / 32;
}

int sm_count() { return vi_sm_count(); }

// Scratch budget for the in-flight eigen-systems when the caller does not size the batch itself: half of the
// device memory that is free right now, at most 32 GiB (VI_SCRATCH_GIB overrides the ceiling).
int64_t scratch_budget_bytes() {
  int64_t ceil_gib = 32;
  if (const char* e = getenv("VI_SCRATCH_GIB")) { long v = atol(e); if (v >= 1) ceil_gib = v; }
  int64_t budget = ceil_gib << 30;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 0) {
    const int64_t half = (int64_t)(free_b / 2);
    if (half < budget) budget = half;
  }
  if (budget < ((int64_t)256 << 20)) budget = (int64_t)256 << 20;
  return budget;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  int v = atoi(e);
  return v >= 1 ? v : dflt;
}

// QL kernel geometry (k_tql_smem): one CTA per SM of W warps, at most Lmax lanes of each carry a system
// (2n doubles of shared memory per system).  Returns 0 warps when not even 32 systems fit.
int ql_fit(int n) { return (int)((227 * 1024) / ((size_t)2 * n * sizeof(double))); }      // systems per SM
int ql_lanes_max(int n = 0) {
  static const int L0 = env_int("VI_TQL_LANES", 12);
  int L = L0 > 32 ? 32 : L0;
  if (n > 0 && L > ql_fit(n)) L = ql_fit(n);      // high orders (N = 500: 29 systems per SM): fewer lanes, never none
  return L;
}
int ql_warps(int n) {
  const int fit = ql_fit(n);
  if (fit < 1) return 0;
  int W = fit / ql_lanes_max(n);
  if (W > 24) W = 24;
  return W < 1 ? 1 : W;
}
int64_t ql_wave(int n) { return (int64_t)sm_count() * ql_warps(n) * ql_lanes_max(n); }

int64_t default_system_cap(int64_t wanted, int n, int nreg, int P) {
  int64_t per = per_system_bytes(n, nreg, P);
  int64_t budget = scratch_budget_bytes();
  int64_t cap = budget / per;
  // whole waves of the QL kernel (thread per system, residency bounded by 2n doubles of shared memory per
  // system): a chunk of 1.1 waves would cost two
  const int64_t wave = ql_wave(n);
  if (wave >= 32 && cap > wave) cap = cap / wave * wave;
  if (cap > wanted) cap = wanted;
  if (cap < 32) cap = 32;
  return vi_align_up(cap, 32);
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t ileave(int64_t s, int n) { return (s >> 5) * 32 * (int64_t)n + (s & 31); }

template <bool GX>
__global__ void __launch_bounds__(1024)
k_tridiag(const double* __restrict__ G, const double* __restrict__ y, const double* __restrict__ regs, SysBuf B,
          Downdate dd) {
  extern __shared__ __align__(16) double sm[];
  const int64_t s = blockIdx.x;
  const int r = B.rec[s];
  if (r < 0) { if (threadIdx.x == 0) B.st[s] = kSkip; return; }
  const int n = B.n, nt = blockDim.x, tid = threadIdx.x;
  vi_tri_ws S;
  double* aux;
  S.ld = B.ld;
  if (GX) { S.X = B.Xg + s * (int64_t)n * S.ld; aux = sm; }
  else { S.X = sm; aux = sm + (size_t)n * S.ld; }      // X stays a provable shared-memory pointer (LDS/STS)
  vi_tri_carve(S, aux, n, nt);
  const double* arow = nullptr;
  double wj = 0.0, bj = 0.0;
  if (dd.A != nullptr) {
    const int j = B.gate[s];
    arow = dd.A + (int64_t)j * n;
    wj = dd.Wm[(int64_t)r * dd.P + j];
    bj = dd.bm[(int64_t)r * dd.P + j];
  }
  vi_tri_load(S, n, G + (int64_t)r * n * n, y + (int64_t)r * n, regs, B.lam + s * (B.nreg > 0 ? B.nreg : 1), B.nreg, tid, nt,
              arow, wj, bj);
  const bool bad = S.sc[1] != 0.0;
  if (!bad) vi_tri_reduce(S, n, B.V + s * B.vstride, tid, nt);
  const int64_t base = ileave(s, n);
  if (!bad)
    for (int i = tid; i < n; i += nt) {
      B.d[base + (int64_t)i * 32] = S.d[i];
      B.e[base + (int64_t)i * 32] = S.e[i];
      B.g[base + (int64_t)i * 32] = S.yv[i];
      B.tau[base + (int64_t)i * 32] = S.tau[i];
    }
  if (tid == 0) { B.scl[s] = S.sc[0]; B.st[s] = bad ? VI_ST_NONFINITE : VI_ST_OK; }
}

// Packed-triangle variant (vi_tridiag_packed.h): 87.5 KB of shared memory at n = 144 -> two CTAs per SM.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_tridiag_packed(const double* __restrict__ G, const double* __restrict__ y, const double* __restrict__ regs, SysBuf B,
                 Downdate dd) {
  extern __shared__ __align__(16) double sm[];
  const int64_t s = blockIdx.x;
  const int r = B.rec[s];
  if (r < 0) { if (threadIdx.x == 0) B.st[s] = kSkip; return; }
  const int n = B.n, nt = blockDim.x, tid = threadIdx.x;
  vi_trp_ws W;
  vi_trp_carve(W, sm, n);
  const double* arow = nullptr;
  double wj = 0.0, bj = 0.0;
  if (dd.A != nullptr) {
    const int j = B.gate[s];
    arow = dd.A + (int64_t)j * n;
    wj = dd.Wm[(int64_t)r * dd.P + j];
    bj = dd.bm[(int64_t)r * dd.P + j];
  }
  vi_trp_load(W, G + (int64_t)r * n * n, y + (int64_t)r * n, regs, B.lam + s * (B.nreg > 0 ? B.nreg : 1), B.nreg, tid, nt,
              arow, wj, bj);
  const bool bad = W.sc[1] != 0.0;
  if (!bad) vi_trp_reduce(W, B.V + s * B.vstride, tid, nt);
  const int64_t base = ileave(s, n);
  if (!bad)
    for (int i = tid; i < n; i += nt) {
      B.d[base + (int64_t)i * 32] = W.d[i];
      B.e[base + (int64_t)i * 32] = W.e[i];
      B.g[base + (int64_t)i * 32] = W.yv[i];
      B.tau[base + (int64_t)i * 32] = W.tau[i];
    }
  if (tid == 0) { B.scl[s] = W.sc[0]; B.st[s] = bad ? VI_ST_NONFINITE : VI_ST_OK; }
}

// ---- two-stage tridiagonalisation (vi_band.h, vi_chase.h) ---------------------------------------------------------
// Stage 1, one CTA per system: X in 8 x 8 blocks in shared memory (87.5 KB at n = 144 + 25 KB of panel factors: two
// CTAs per SM), reduced to a band of half-width 8 by 17 block-reflector steps whose O(n^3) work is DMMA.  Writes the
// band and g = Q1^T y to B.band, T and V of every panel to B.V.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_band(const double* __restrict__ G, const double* __restrict__ y, const double* __restrict__ regs, SysBuf B, Downdate dd) {
  extern __shared__ __align__(16) double sm[];
  const int64_t s = blockIdx.x;
  const int r = B.rec[s];
  if (r < 0) { if (threadIdx.x == 0) B.st[s] = kSkip; return; }
  const int n = B.n;
  vi_bnd_ws W;
  vi_bnd_carve(W, sm, n);
  const double* arow = nullptr;
  double wj = 0.0, bj = 0.0;
  if (dd.A != nullptr) {
    const int j = B.gate[s];
    arow = dd.A + (int64_t)j * n;
    wj = dd.Wm[(int64_t)r * dd.P + j];
    bj = dd.bm[(int64_t)r * dd.P + j];
  }
  vi_bnd_load(W, G + (int64_t)r * n * n, y + (int64_t)r * n, regs, B.lam + s * (B.nreg > 0 ? B.nreg : 1), B.nreg, arow, wj, bj);
  const bool bad = W.sc[1] != 0.0;
  if (!bad) {
    double* band = B.band + s * (int64_t)vi_bnd_band_doubles(n);
    if (B.split_p1 > 0) {
      // first kernel of the split reduction: the trailing matrix goes to the (still unused) tape area of the system
      vi_bnd_reduce(W, B.V + s * B.vstride, B.split_p1);
      vi_bnd_store_band(W, band, band + 9 * W.npad, 8 * B.split_p1);
      vi_bnd_store_trailing(W, B.split_p1, B.tcs + s * (int64_t)B.tapecap * 2);
    } else {
      vi_bnd_reduce(W, B.V + s * B.vstride);
      vi_bnd_store_band(W, band);
    }
  }
  if (threadIdx.x == 0) { B.scl[s] = W.sc[0]; B.st[s] = bad ? VI_ST_NONFINITE : VI_ST_OK; }
}

// Stage 1 for orders whose X does not fit shared memory: blocks in global memory (B.Xg), panel factors in shared memory.
__global__ void __launch_bounds__(kBigWarps * 32, 2)
k_band_big(const double* __restrict__ G, const double* __restrict__ y, const double* __restrict__ regs, SysBuf B, Downdate dd) {
  extern __shared__ __align__(16) double sm[];
  const int64_t s = blockIdx.x;
  const int r = B.rec[s];
  if (r < 0) { if (threadIdx.x == 0) B.st[s] = kSkip; return; }
  const int n = B.n;
  vi_bnd_ws W;
  vi_bnd_carve_big(W, sm, B.Xg + s * (int64_t)n * B.ld, n, kBigWarps);
  const double* arow = nullptr;
  double wj = 0.0, bj = 0.0;
  if (dd.A != nullptr) {
    const int j = B.gate[s];
    arow = dd.A + (int64_t)j * n;
    wj = dd.Wm[(int64_t)r * dd.P + j];
    bj = dd.bm[(int64_t)r * dd.P + j];
  }
  vi_bnd_load(W, G + (int64_t)r * n * n, y + (int64_t)r * n, regs, B.lam + s * (B.nreg > 0 ? B.nreg : 1), B.nreg, arow, wj, bj);
  const bool bad = W.sc[1] != 0.0;
  if (!bad) {
    vi_bnd_reduce_big(W, B.V + s * B.vstride);
    vi_bnd_store_band(W, B.band + s * (int64_t)vi_bnd_band_doubles(n));
  }
  if (threadIdx.x == 0) { B.scl[s] = W.sc[0]; B.st[s] = bad ? VI_ST_NONFINITE : VI_ST_OK; }
}

// Second kernel of the split stage 1: the trailing matrix of order n - 8 p1 as an independent band reduction
// (panel q of it = panel p1 + q of the system), four CTAs of four warps per SM.
__global__ void __launch_bounds__(kTailWarps * 32, 4)
k_band_tail(SysBuf B) {
  extern __shared__ __align__(16) double sm[];
  const int64_t s = blockIdx.x;
  if (B.st[s] != VI_ST_OK) return;
  const int n = B.n, p1 = B.split_p1, np = vi_bnd_npad(n);
  vi_bnd_ws W;
  vi_bnd_carve(W, sm, n - 8 * p1, kTailWarps);
  vi_bnd_load_trailing(W, B.tcs + s * (int64_t)B.tapecap * 2);
  vi_bnd_reduce<3>(W, B.V + s * B.vstride + vi_bnd_voff(np, p1));
  double* band = B.band + s * (int64_t)vi_bnd_band_doubles(n);
  vi_bnd_store_band(W, band + 9 * 8 * p1, band + 9 * np + 8 * p1, W.npad);
}

// Stage 2, one WARP per system: band -> tridiagonal by bulge chasing, four sweeps in flight (one per 8 lanes).
// Leaves d, e, g = Q2^T Q1^T y in the interleaved layout the QL kernels read; reflectors behind stage 1's in B.V.
constexpr int kChaseWarps = 5;     // 5 x 21.5 KB at n = 144: two CTAs = 10 systems per SM
__global__ void __launch_bounds__(kChaseWarps * 32)
k_chase(int64_t nsys, SysBuf B) {
  extern __shared__ __align__(16) double sm[];
  const int n = B.n, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;       // (high orders: fewer warps per CTA)
  if (s >= nsys || B.st[s] != VI_ST_OK) return;
  const int np = vi_bnd_npad(n);
  double* Bw = sm + (size_t)warp * vi_chs_doubles(n);
  double* g = Bw + VI_CHS_LDB * np;
  vi_chs_load(Bw, g, B.band + s * (int64_t)vi_bnd_band_doubles(n), n);
  vi_chs_reduce(Bw, g, n, B.V + s * B.vstride + vi_bnd_vdoubles(n));
  const int64_t base = ileave(s, n);
  for (int i = lane; i < n; i += 32) {
    B.d[base + (int64_t)i * 32] = Bw[i * VI_CHS_LDB];
    B.e[base + (int64_t)i * 32] = (i + 1 < n) ? Bw[i * VI_CHS_LDB + 1] : 0.0;
    B.g[base + (int64_t)i * 32] = g[i];
  }
}

// c = Q1 Q2 u, one warp per system (two-stage counterpart of k_apply)
__global__ void __launch_bounds__(kChaseWarps * 32)
k_apply_ts(int64_t nsys, SysBuf B, double* __restrict__ Cout, int32_t* __restrict__ rank_out) {
  extern __shared__ __align__(16) double sm[];
  const int n = B.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * kChaseWarps + warp;
  if (s >= nsys) return;
  const int st = B.st[s];
  if (st == kSkip) return;
  double* Cs = Cout + s * (int64_t)n;
  if (st != VI_ST_OK) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = lane; i < n; i += 32) Cs[i] = nan;
    if (lane == 0) rank_out[s] = 0;
    return;
  }
  const int np = vi_bnd_npad(n);
  double* u = sm + (size_t)warp * (np + 8);
  const int64_t base = ileave(s, n);
  for (int i = lane; i < np + 8; i += 32) u[i] = (i < n) ? B.g[base + (int64_t)i * 32] : 0.0;
  __syncwarp();
  const double* Vg = B.V + s * B.vstride;
  vi_chs_apply_q(u, n, Vg + vi_bnd_vdoubles(n));
  vi_bnd_apply_q(u, n, Vg);
  for (int i = lane; i < n; i += 32) Cs[i] = u[i];
  if (lane == 0) rank_out[s] = B.rank[s];
}

__device__ __forceinline__ vi_tape tape_of(const SysBuf& B, int64_t s) {
  double* cs = B.tcs + s * (int64_t)B.tapecap * 2;
  return vi_tape{{cs, 2}, {cs + 1, 2}, {B.tix + s * (int64_t)B.tapecap, 1}, B.tapecap};
}

// Tape replay, one THREAD per system (every lane does distinct work; a warp-per-system version spent
// 26 warp-instructions per rotation on redundant lanes).  The lane's vector lives in shared memory laid
// out [i][lane]; its tape is contiguous in global memory and is fetched four entries ahead.
// DIR = +1: w <- Z^T w (tape order), DIR = -1: w <- Z w (reverse order).  Lanes whose tape is shorter
// than the longest in the warp apply identity rotations.
template <int DIR>
__device__ __forceinline__ void tape_replay(vi_svec w, const double2* __restrict__ cs,
                                            const int32_t* __restrict__ ix, int32_t nrot, int32_t nmax) {
  constexpr int NB = 8;
  double2 cur[NB], nxt[NB];
  int32_t icur[NB], inxt[NB];
  auto fetch = [&](int32_t base, double2 (&c)[NB], int32_t (&id)[NB]) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int32_t t = base + q;
      const bool ok = t < nrot;
      const int32_t tt = ok ? ((DIR > 0) ? t : nrot - 1 - t) : 0;
      const double2 v = cs[tt];
      const int32_t code = ix[tt];
      c[q] = ok ? v : make_double2(1.0, 0.0);
      id[q] = ok ? code : 0;
    }
  };
  fetch(0, cur, icur);
  for (int32_t base = 0; base < nmax; base += NB) {
    fetch(base + NB, nxt, inxt);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int pi = icur[q] >> 1;
      const int pj = (icur[q] & 1) ? pi - 1 : pi + 1;
      const double c = cur[q].x, sn = cur[q].y;
      const double a = w[pi], bb = w[pj];
      if (DIR > 0) { w[pj] = sn * a + c * bb; w[pi] = c * a - sn * bb; }
      else { w[pi] = c * a + sn * bb; w[pj] = c * bb - sn * a; }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) { cur[q] = nxt[q]; icur[q] = inxt[q]; }
  }
}

// u = Z L^+ Z^T g, one thread per system: forward replay, spectral cut-off (|l| > rcond max|l|: gelsd's
// rule), backward replay.  Result (scaled back by 2^-ex) returns to B.g; rank to B.rank.
// Both thread-per-system kernels are latency-bound dependent chains whose residency is capped by shared
// memory (n doubles per system here, 2n in k_tql_smem), not by threads.  They therefore run with only L of the
// 32 lanes of a warp carrying a system ("sparse lanes"): the same number of systems per SM is spread over
// 32/L times as many warps, which gives the schedulers that many more independent chains to interleave and
// cuts the divergence between the systems of a warp.  L adapts to the batch (small Brent rounds: L = 1).
constexpr int kReplayWarpsMax = 12;
__global__ void __launch_bounds__(kReplayWarpsMax * 32)
k_replay(int64_t nsys, SysBuf B, double rcond, int L) {
  const int kReplayWarps = blockDim.x >> 5;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = B.n;
  const int T = kReplayWarps * L;                       // systems per CTA
  const int slot = warp * L + (lane < L ? lane : 0);
  const int64_t s = (int64_t)blockIdx.x * T + slot;
  const bool act = (lane < L) && (s < nsys) && (B.st[s] == VI_ST_OK);
  vi_svec w{sm + slot, T};
  const int64_t sc = act ? s : 0;
  const int64_t base = ileave(sc, n);
  const int32_t nrot = act ? B.nrot[sc] : 0;
  int32_t nmax = nrot;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  if (nmax == 0 && !__any_sync(0xffffffffu, act)) return;
  if (act)
    for (int i = 0; i < n; ++i) w[i] = B.g[base + (int64_t)i * 32];
  const double2* tcs = reinterpret_cast<const double2*>(B.tcs + sc * (int64_t)B.tapecap * 2);
  const int32_t* tix = B.tix + sc * (int64_t)B.tapecap;
  if (lane < L) tape_replay<+1>(w, tcs, tix, nrot, nmax);
  if (act) {
    double lmax = 0.0;
    for (int i = 0; i < n; ++i) lmax = fmax(lmax, fabs(B.d[base + (int64_t)i * 32]));
    const double cut = rcond * lmax;
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      const double l = B.d[base + (int64_t)i * 32];
      if (fabs(l) > cut) { w[i] = w[i] / l; ++rank; }
      else w[i] = 0.0;
    }
    B.rank[s] = rank;
  }
  if (lane < L) tape_replay<-1>(w, tcs, tix, nrot, nmax);
  if (act) {
    const double scl = B.scl[s];
    for (int i = 0; i < n; ++i) B.g[base + (int64_t)i * 32] = w[i] * scl;
  }
}

// QL proper, one THREAD per system: eigenvalues + rotation tape (vi_tql_values).  d and e (2n doubles)
// sit on the dependency chain of every rotation, so they live in shared memory laid out [i][thread]
// (a lane always hits its own bank pair whatever i it is at).  Leaves the eigenvalues in B.d, the
// tape in B.tcs / B.tix; the right-hand side is handled by k_apply.
__global__ void k_tql_smem(int64_t nsys, SysBuf B, int L, int iter_batch) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = B.n;
  const int T = (blockDim.x >> 5) * L;                  // systems per CTA (sparse lanes, see k_replay)
  const int slot = warp * L + (lane < L ? lane : 0);
  const int64_t s = (int64_t)blockIdx.x * T + slot;
  const bool act = (lane < L) && (s < nsys) && (B.st[s] == VI_ST_OK);     // idle lanes still take part in the warp votes
  const int64_t sc = act ? s : 0;
  vi_svec d{sm + slot, T}, e{sm + (size_t)n * T + slot, T};
  const int64_t base = ileave(sc, n);
  if (act)
    for (int i = 0; i < n; ++i) {
      d[i] = B.d[base + (int64_t)i * 32];
      e[i] = B.e[base + (int64_t)i * 32];
    }
  int32_t nrot = 0;
  const int q = vi_tql_values_flat(n, d, e, tape_of(B, sc), &nrot, act, iter_batch);
  if (!act) return;
  if (q != 0) { B.st[s] = VI_ST_NOCONV; return; }
  B.nrot[s] = nrot;
  if (B.rot_total) atomicAdd(B.rot_total, (unsigned long long)nrot);
  for (int i = 0; i < n; ++i) B.d[base + (int64_t)i * 32] = d[i];
}

// One system per warp (small Brent rounds, L = 1): lane 0 runs the plain nested-loop QL, which has less
// per-rotation overhead than the per-lane state machine (that one only pays when several lanes share a warp).
// Same arithmetic, bit-identical eigenvalues and tape (tests/test_host_math.py).
__global__ void k_tql_single(int64_t nsys, SysBuf B) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = B.n;
  const int T = blockDim.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * T + warp;
  if (lane != 0 || s >= nsys || B.st[s] != VI_ST_OK) return;
  vi_svec d{sm + warp, T}, e{sm + (size_t)n * T + warp, T};
  const int64_t base = ileave(s, n);
  for (int i = 0; i < n; ++i) {
    d[i] = B.d[base + (int64_t)i * 32];
    e[i] = B.e[base + (int64_t)i * 32];
  }
  int32_t nrot = 0;
  const int q = vi_tql_values(n, d, e, tape_of(B, s), &nrot);
  if (q != 0) { B.st[s] = VI_ST_NOCONV; return; }
  B.nrot[s] = nrot;
  if (B.rot_total) atomicAdd(B.rot_total, (unsigned long long)nrot);
  for (int i = 0; i < n; ++i) B.d[base + (int64_t)i * 32] = d[i];
}

// c = Q u, one WARP per system: the reflectors are applied cooperatively (coalesced V rows, shuffle-tree
// dot products) to the vector k_replay left in B.g.
constexpr int kApplyWarps = 8;
__global__ void __launch_bounds__(kApplyWarps * 32)
k_apply(int64_t nsys, SysBuf B, double* __restrict__ Cout, int32_t* __restrict__ rank_out) {
  extern __shared__ __align__(16) double sm[];
  const int n = B.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * kApplyWarps + warp;
  if (s >= nsys) return;
  const int st = B.st[s];
  if (st == kSkip) return;
  double* Cs = Cout + s * (int64_t)n;
  if (st != VI_ST_OK) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = lane; i < n; i += 32) Cs[i] = nan;
    if (lane == 0) rank_out[s] = 0;
    return;
  }
  double* w = sm + (size_t)warp * n;
  const int64_t base = ileave(s, n);
  for (int i = lane; i < n; i += 32) w[i] = B.g[base + (int64_t)i * 32];
  __syncwarp();
  const double* V = B.V + s * B.vstride;
  for (int j = n - 3; j >= 0; --j) {
    const double t = B.tau[base + (int64_t)j * 32];
    if (t == 0.0) continue;
    const double* vj = V + (int64_t)j * n;
    double dot = 0.0;
    for (int i = j + 1 + lane; i < n; i += 32) dot += vj[i] * w[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    dot *= t;
    for (int i = j + 1 + lane; i < n; i += 32) w[i] -= dot * vj[i];
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) Cs[i] = w[i];
  if (lane == 0) rank_out[s] = B.rank[s];
}

// u = Z L^+ Z^T g and c = Q u in ONE kernel, one WARP per system: the rotation tape is replayed as a wavefront of
// QL sweeps across the lanes (vi_wave.h: ~7 rotations per step on average at n = 144 instead of one), the spectral
// cut-off (|l| > rcond max|l|: gelsd's rule) runs lane-parallel, and the back-transformation (two-stage: chase
// reflectors then block reflectors; one-stage: Householder rows) follows on the vector still in shared memory.
// Replaces k_replay + k_apply / k_apply_ts (kept behind VI_OLD_REPLAY for A/B measurements).
constexpr int kWaveWarps = 8;
__global__ void __launch_bounds__(kWaveWarps * 32)
k_replay_wave(int64_t nsys, SysBuf B, double rcond, double* __restrict__ Cout, int32_t* __restrict__ rank_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int n = B.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * kWaveWarps + warp;
  if (s >= nsys) return;
  const int st = B.st[s];
  if (st == kSkip) return;
  double* Cs = Cout + s * (int64_t)n;
  if (st != VI_ST_OK) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = lane; i < n; i += 32) Cs[i] = nan;
    if (lane == 0) rank_out[s] = 0;
    return;
  }
  const int np = (n + 7) & ~7;
  // the sweep table lives in global memory (B.wtab) when there is one: shared memory then holds only the vector, and
  // the SM keeps most of its 256 KB as L1 for the 32 tape streams of every warp
  unsigned char* mine = smraw + (size_t)warp * (B.wtab ? vi_wav_bytes_vec(n) : vi_wav_bytes(n));
  double* w = reinterpret_cast<double*>(mine);
  int32_t* tab = B.wtab ? B.wtab + s * (int64_t)(vi_wav_maxsweeps(n) + 2) : reinterpret_cast<int32_t*>(mine + np * 8 + 64);
  const int64_t base = ileave(s, n);
  for (int i = lane; i < np + 8; i += 32) w[i] = (i < n) ? B.g[base + (int64_t)i * 32] : 0.0;
  const int32_t nrot = B.nrot[s];
  const double* cs = B.tcs + s * (int64_t)B.tapecap * 2;
  const int32_t* ix = B.tix + s * (int64_t)B.tapecap;
  int tnext = 0;
  const int ns = vi_wav_scan(ix, 0, nrot, tab, vi_wav_maxsweeps(n), &tnext);
  const bool whole = tnext == nrot;            // the sweep table holds the whole tape (else: sequential replay, rare)
  if (whole) vi_wav_pass<true>(w, cs, ix, tab, ns);
  else if (lane == 0) vi_tape_apply_zt(vi_svec{w, 1}, vi_tape{{const_cast<double*>(cs), 2}, {const_cast<double*>(cs) + 1, 2}, {const_cast<int32_t*>(ix), 1}, B.tapecap}, nrot);
  __syncwarp();
  // spectral cut-off
  double lmax = 0.0;
  for (int i = lane; i < n; i += 32) lmax = fmax(lmax, fabs(B.d[base + (int64_t)i * 32]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  const double cut = rcond * lmax;
  int rank = 0;
  for (int i = lane; i < n; i += 32) {
    const double l = B.d[base + (int64_t)i * 32];
    if (fabs(l) > cut) { w[i] = w[i] / l; ++rank; }
    else w[i] = 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  __syncwarp();
  if (whole) vi_wav_pass<false>(w, cs, ix, tab, ns);
  else if (lane == 0) vi_tape_apply_z(vi_svec{w, 1}, vi_tape{{const_cast<double*>(cs), 2}, {const_cast<double*>(cs) + 1, 2}, {const_cast<int32_t*>(ix), 1}, B.tapecap}, nrot);
  __syncwarp();
  const double scl = B.scl[s];
  for (int i = lane; i < n; i += 32) w[i] *= scl;
  __syncwarp();
  const double* Vg = B.V + s * B.vstride;
  if (B.two_stage) {
    vi_chs_apply_q(w, n, Vg + vi_bnd_vdoubles(n));
    vi_bnd_apply_q(w, n, Vg);
  } else {
    for (int j = n - 3; j >= 0; --j) {           // Householder rows of the one-stage reduction (as k_apply)
      const double t = B.tau[base + (int64_t)j * 32];
      if (t == 0.0) continue;
      const double* vj = Vg + (int64_t)j * n;
      double dot = 0.0;
      for (int i = j + 1 + lane; i < n; i += 32) dot += vj[i] * w[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dot *= t;
      for (int i = j + 1 + lane; i < n; i += 32) w[i] -= dot * vj[i];
      __syncwarp();
    }
  }
  for (int i = lane; i < n; i += 32) Cs[i] = w[i];
  if (lane == 0) { rank_out[s] = rank; B.rank[s] = rank; }
}

// The same for FULL batches, four systems per warp (lane groups of 8, vi_wav_pass<., 8>): the sweeps of a QL run are
// short on average, so with one system per warp ~7 of 32 lanes rotate per step and the kernel is bound by the
// instruction issue of the idle ones.  The scans, the spectral cut-off and the back-transformations use the whole
// warp, one system after the other; the sweep tables live in global memory (B.wtab) so that shared memory holds
// nothing but the four vectors.
constexpr int kWaveSys = 4;
__global__ void __launch_bounds__(kWaveWarps * 32)
k_replay_wave4(int64_t nsys, SysBuf B, double rcond, double* __restrict__ Cout, int32_t* __restrict__ rank_out) {
  extern __shared__ __align__(16) double smw[];
  const int n = B.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 3;
  const int np = (n + 7) & ~7;
  const int64_t s0 = ((int64_t)blockIdx.x * kWaveWarps + warp) * kWaveSys;
  if (s0 >= nsys) return;
  double* wbase = smw + (size_t)warp * kWaveSys * (np + 8);
  const int maxsw = vi_wav_maxsweeps(n);
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  int my_ns = 0;                                 // sweeps of this lane group's system (0: nothing to replay in parallel)
  unsigned okmask = 0, seqmask = 0;              // systems with a solution to compute / whose table overflowed
#pragma unroll 1
  for (int g = 0; g < kWaveSys; ++g) {
    const int64_t s = s0 + g;
    if (s >= nsys) break;
    const int st = B.st[s];
    if (st == kSkip) continue;
    double* Cs = Cout + s * (int64_t)n;
    if (st != VI_ST_OK) {
      for (int i = lane; i < n; i += 32) Cs[i] = nan;
      if (lane == 0) rank_out[s] = 0;
      continue;
    }
    okmask |= 1u << g;
    double* w = wbase + g * (np + 8);
    const int64_t base = ileave(s, n);
    for (int i = lane; i < np + 8; i += 32) w[i] = (i < n) ? B.g[base + (int64_t)i * 32] : 0.0;
    int tnext = 0;
    const int32_t nrot = B.nrot[s];
    const int k = vi_wav_scan(B.tix + s * (int64_t)B.tapecap, 0, nrot, B.wtab + s * (int64_t)(maxsw + 2), maxsw, &tnext);
    if (tnext != nrot) seqmask |= 1u << g;
    else if (g == grp) my_ns = k;
  }
  __syncwarp();
  const int64_t sg = s0 + grp;                   // this lane group's system
  const bool mine = ((okmask >> grp) & 1u) != 0;
  const int64_t sc = mine ? sg : s0;
  double* wg = wbase + grp * (np + 8);
  const double* csg = B.tcs + sc * (int64_t)B.tapecap * 2;
  const int32_t* ixg = B.tix + sc * (int64_t)B.tapecap;
  const int32_t* tabg = B.wtab + sc * (int64_t)(maxsw + 2);
  auto sequential = [&](bool fwd) {              // table overflow (never seen on this path's spectra): one lane
    for (int g = 0; g < kWaveSys; ++g) {
      if (!((seqmask >> g) & 1u) || lane != 0) continue;
      const int64_t s = s0 + g;
      double* cs = B.tcs + s * (int64_t)B.tapecap * 2;
      vi_tape tp{{cs, 2}, {cs + 1, 2}, {B.tix + s * (int64_t)B.tapecap, 1}, B.tapecap};
      if (fwd) vi_tape_apply_zt(vi_svec{wbase + g * (np + 8), 1}, tp, B.nrot[s]);
      else vi_tape_apply_z(vi_svec{wbase + g * (np + 8), 1}, tp, B.nrot[s]);
    }
    __syncwarp();
  };
  vi_wav_pass<true, 8>(wg, csg, ixg, tabg, mine ? my_ns : 0);
  if (seqmask) sequential(true);
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < kWaveSys; ++g) {
    if (!((okmask >> g) & 1u)) continue;
    const int64_t s = s0 + g;
    double* w = wbase + g * (np + 8);
    const int64_t base = ileave(s, n);
    double lmax = 0.0;
    for (int i = lane; i < n; i += 32) lmax = fmax(lmax, fabs(B.d[base + (int64_t)i * 32]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    const double cut = rcond * lmax;
    int rank = 0;
    for (int i = lane; i < n; i += 32) {
      const double l = B.d[base + (int64_t)i * 32];
      if (fabs(l) > cut) { w[i] = w[i] / l; ++rank; }
      else w[i] = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) { rank_out[s] = rank; B.rank[s] = rank; }
  }
  __syncwarp();
  vi_wav_pass<false, 8>(wg, csg, ixg, tabg, mine ? my_ns : 0);
  if (seqmask) sequential(false);
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < kWaveSys; ++g) {
    if (!((okmask >> g) & 1u)) continue;
    const int64_t s = s0 + g;
    double* w = wbase + g * (np + 8);
    const double scl = B.scl[s];
    for (int i = lane; i < n; i += 32) w[i] *= scl;
    __syncwarp();
    const double* Vg = B.V + s * B.vstride;
    vi_chs_apply_q(w, n, Vg + vi_bnd_vdoubles(n));
    vi_bnd_apply_q(w, n, Vg);
    double* Cs = Cout + s * (int64_t)n;
    for (int i = lane; i < n; i += 32) Cs[i] = w[i];
  }
}

// ---- covariance: dC = H (A^T W A) H, H = pinv(X)  (interpolate.py:464-467) ---------------------
// scipy.linalg.pinv keeps singular values above max(M,N)*eps*s_max; for the symmetrised X that is
// |eigenvalue| > N*eps*max|eigenvalue|.  The eigenvectors E = Q Z are formed once per record:
// thread i owns column i of Z in shared memory and replays the rotation tape on it (all threads run
// the same rotation sequence: uniform control flow, tape entries are broadcast loads), then applies the
// reflectors.  H = E diag(scl/lambda) E^T, T = H G and dC = T H are three batched N x N products.
// The rotation tape and the reflectors are the same for all n columns, so they are STAGED through shared memory
// (cp.async, double-buffered: 256 rotations / 8 reflector rows at a time) instead of every thread fetching every
// entry from global memory behind its own dependent chain (that version spent 2.2 ms per record, 95 % of the
// covariance phase).
constexpr int kEvTape = 256;     // rotations per staged chunk
constexpr int kEvRows = 8;       // reflector rows per staged batch

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cpa8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cpa4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_)); }

// doubles of staging space behind the n x ld eigenvector block
__host__ __device__ inline int eigvec_stage_doubles(int n) {
  const int tape = 2 * kEvTape * 2 + (2 * kEvTape + 1) / 2;                  // (c, s) pairs + int32 codes, two buffers
  const int rows = 2 * kEvRows * ((n + 1) & ~1) + 2 * kEvRows;               // rows + their tau, two buffers
  const int npad = (n + 7) & ~7;
  const int pan = 2 * (64 + 8 * (npad > 8 ? npad - 8 : 0));                  // two-stage: T + V of a panel, two buffers
  const int refl = 2 * 128 * 8;                                              // two-stage: 128 reflectors, two buffers
  int m = tape > rows ? tape : rows;
  if (pan > m) m = pan;
  if (refl > m) m = refl;
  return m + 2;
}

// GLOBAL = true: orders whose n x ld block does not fit shared memory keep the eigenvector block in its output
// array (row-major, thread i owns column i: every access is coalesced across the CTA); only the staging buffers
// live in shared memory.  Functional path for large models (N > VI_NMAX_SMEM), not a tuned one.
template <bool GLOBAL>
__global__ void k_eigvec(int64_t s0, SysBuf B, double pinv_rtol, double* __restrict__ E, double* __restrict__ dinv) {
  extern __shared__ __align__(16) double sm[];
  const int n = B.n, i = threadIdx.x, ld = GLOBAL ? B.n : B.ld, nt = blockDim.x;
  const int64_t s = s0 + blockIdx.x;
  if (B.st[s] != VI_ST_OK || B.rec[s] < 0) return;
  const int64_t base = ileave(s, n);
  double* stage = GLOBAL ? sm : sm + (size_t)n * ld;
  double* col = (GLOBAL ? E + (int64_t)blockIdx.x * n * n : sm) + i;      // this thread's column of Z (stride ld)
  if (i < n)
    for (int r = 0; r < n; ++r) col[r * ld] = (r == i) ? 1.0 : 0.0;
  // ---- Z: replay the tape backwards (vi_tape_apply_z order) ------------------------------------------------
  {
    const int32_t nrot = B.nrot[s];
    const double2* gcs = reinterpret_cast<const double2*>(B.tcs + s * (int64_t)B.tapecap * 2);
    const int32_t* gix = B.tix + s * (int64_t)B.tapecap;
    double2* scs = reinterpret_cast<double2*>(stage);                  // [2][kEvTape]
    int32_t* six = reinterpret_cast<int32_t*>(scs + 2 * kEvTape);      // [2][kEvTape]
    const int nchunk = (nrot + kEvTape - 1) / kEvTape;
    auto load = [&](int c) {
      const int buf = c & 1;
      for (int t = i; t < kEvTape; t += nt) {
        const int32_t idx = nrot - 1 - (c * kEvTape + t);              // t-th rotation of this chunk, going backwards
        if (idx >= 0) { cpa16(scs + buf * kEvTape + t, gcs + idx); cpa4(six + buf * kEvTape + t, gix + idx); }
      }
      cpa_commit();
    };
    if (nchunk > 0) load(0);
    for (int c = 0; c < nchunk; ++c) {
      if (c + 1 < nchunk) { load(c + 1); cpa_wait<1>(); } else cpa_wait<0>();
      __syncthreads();
      if (i < n) {
        const int cnt = min(kEvTape, nrot - c * kEvTape);
        const double2* pcs = scs + (c & 1) * kEvTape;
        const int32_t* pix = six + (c & 1) * kEvTape;
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
          const int32_t code = pix[t];
          const double2 cs = pcs[t];
          const int pi = code >> 1;
          const int pj = (code & 1) ? pi - 1 : pi + 1;
          const double a = col[pi * ld], b = col[pj * ld];
          col[pi * ld] = cs.x * a + cs.y * b;
          col[pj * ld] = cs.x * b - cs.y * a;
        }
      }
      __syncthreads();
    }
  }
  // ---- E = Q1 Q2 Z (two-stage reduction): stage-2 reflectors in reverse order, then the block reflectors -------
  if (B.two_stage) {
    const double* Vg = B.V + s * B.vstride;
    {
      constexpr int CH = 128;                                 // reflectors per staged chunk (8 doubles each)
      const double* R = Vg + vi_bnd_vdoubles(n);
      const int nrefl = vi_chs_nrefl(n);
      const int nchunk = (nrefl + CH - 1) / CH;
      auto load = [&](int c) {                                // chunk c = reflectors [hi - CH, hi), hi = nrefl - c CH
        const int hi = nrefl - c * CH, lo = hi - CH > 0 ? hi - CH : 0;
        double* dst = stage + (c & 1) * CH * 8;
        for (int e = i; e < (hi - lo) * 4; e += nt) cpa16(dst + 2 * e, R + (int64_t)lo * 8 + 2 * e);
        cpa_commit();
      };
      int sw = vi_chs_nsweeps(n) - 1;                         // (sweep, step) of the reflector being applied
      int kk = sw >= 0 ? vi_chs_nsteps(n, sw) - 1 : 0;
      if (nchunk > 0) load(0);
      for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) { load(c + 1); cpa_wait<1>(); } else cpa_wait<0>();
        __syncthreads();
        const int hi = nrefl - c * CH, lo = hi - CH > 0 ? hi - CH : 0;
        const double* src = stage + (c & 1) * CH * 8;
        for (int idx = hi - 1; idx >= lo; --idx) {
          const double* rv = src + (idx - lo) * 8;
          const double tau = rv[0];
          const int r0 = sw + 1 + 8 * kk;
          if (i < n && tau != 0.0) {
            const int L = (n - r0 < 8) ? n - r0 : 8;
            double x[8], dot = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              x[q] = (q < L) ? col[(r0 + q) * ld] : 0.0;
              dot = fma((q == 0) ? 1.0 : rv[q], x[q], dot);
            }
            dot *= tau;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < L) col[(r0 + q) * ld] = fma(-dot, (q == 0) ? 1.0 : rv[q], x[q]);
          }
          if (--kk < 0) { --sw; kk = sw >= 0 ? vi_chs_nsteps(n, sw) - 1 : 0; }
        }
        __syncthreads();
      }
    }
    {
      const int npad = vi_bnd_npad(n), nbk = npad >> 3;
      const int npan = nbk - 1;
      const int slot = 64 + 8 * (npad - 8);                   // doubles of the largest panel (T + V)
      auto load = [&](int q) {                                // q-th panel from the end: p = npan - 1 - q
        const int pp = npan - 1 - q;
        const int cnt = 64 + 8 * (npad - 8 * (pp + 1));
        double* dst = stage + (q & 1) * slot;
        const double* srcp = Vg + vi_bnd_voff(npad, pp);
        for (int e = i; e < cnt / 2; e += nt) cpa16(dst + 2 * e, srcp + 2 * e);
        cpa_commit();
      };
      if (npan > 0) load(0);
      for (int q = 0; q < npan; ++q) {
        if (q + 1 < npan) { load(q + 1); cpa_wait<1>(); } else cpa_wait<0>();
        __syncthreads();
        const int pp = npan - 1 - q;
        const int r0 = 8 * (pp + 1), m = npad - r0;
        const int mv = (n - r0 < m) ? n - r0 : m;             // rows that exist in the n x n matrix
        const double* Tm = stage + (q & 1) * slot;
        const double* Vm = Tm + 64;
        if (i < n) {
          double t[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) t[c] = 0.0;
          for (int r = 0; r < mv; ++r) {
            const double x = col[(r0 + r) * ld];
#pragma unroll
            for (int c = 0; c < 8; ++c) t[c] = fma(Vm[c * m + r], x, t[c]);
          }
          double t2[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int c = a; c < 8; ++c) acc = fma(Tm[a * 8 + c], t[c], acc);
            t2[a] = acc;
          }
          for (int r = 0; r < mv; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc = fma(Vm[c * m + r], t2[c], acc);
            col[(r0 + r) * ld] -= acc;
          }
        }
        __syncthreads();
      }
    }
  } else
  // ---- E = Q Z: reflectors n-3 .. 0 (vi_tri_backtransform order), rows staged kEvRows at a time ----------------
  {
    const int nld = (n + 1) & ~1;
    double* srow = stage;                                   // [2][kEvRows][nld]
    double* stau = stage + 2 * kEvRows * nld;               // [2][kEvRows]
    const double* V = B.V + s * B.vstride;
    const int nref = n - 2;                                 // reflectors j = 0 .. n-3
    const int nbatch = nref > 0 ? (nref + kEvRows - 1) / kEvRows : 0;
    auto load = [&](int bidx) {
      const int buf = bidx & 1;
      const int jhi = nref - 1 - bidx * kEvRows;            // first (highest) reflector of the batch
      for (int e = i; e < kEvRows * n; e += nt) {
        const int q = e / n, c = e - q * n;
        const int j = jhi - q;
        if (j >= 0 && c > j) cpa8(srow + (buf * kEvRows + q) * nld + c, V + (int64_t)j * n + c);
      }
      if (i < kEvRows && jhi - i >= 0) cpa8(stau + buf * kEvRows + i, B.tau + base + (int64_t)(jhi - i) * 32);
      cpa_commit();
    };
    if (nbatch > 0) load(0);
    for (int bidx = 0; bidx < nbatch; ++bidx) {
      if (bidx + 1 < nbatch) { load(bidx + 1); cpa_wait<1>(); } else cpa_wait<0>();
      __syncthreads();
      if (i < n) {
        const int jhi = nref - 1 - bidx * kEvRows;
        for (int q = 0; q < kEvRows; ++q) {
          const int j = jhi - q;
          if (j < 0) break;
          const double t = stau[(bidx & 1) * kEvRows + q];
          if (t == 0.0) continue;
          const double* vj = srow + ((bidx & 1) * kEvRows + q) * nld;
          double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
          int r = j + 1;
          for (; r + 3 < n; r += 4) {
            d0 += vj[r] * col[r * ld];
            d1 += vj[r + 1] * col[(r + 1) * ld];
            d2 += vj[r + 2] * col[(r + 2) * ld];
            d3 += vj[r + 3] * col[(r + 3) * ld];
          }
          for (; r < n; ++r) d0 += vj[r] * col[r * ld];
          const double dot = ((d0 + d1) + (d2 + d3)) * t;
          for (r = j + 1; r < n; ++r) col[r * ld] -= dot * vj[r];
        }
      }
      __syncthreads();
    }
  }
  if (i < n) {
    if (!GLOBAL) {
      double* Es = E + (int64_t)blockIdx.x * n * n;
      for (int r = 0; r < n; ++r) Es[(int64_t)r * n + i] = col[r * ld];
    }
    double lmax = 0.0;
    for (int m = 0; m < n; ++m) lmax = fmax(lmax, fabs(B.d[base + (int64_t)m * 32]));
    const double l = B.d[base + (int64_t)i * 32];
    dinv[(int64_t)blockIdx.x * n + i] = (fabs(l) > pinv_rtol * lmax) ? B.scl[s] / l : 0.0;
  }
}

// C[b] = A[b] diag(sc[b]) op(B[b]), all n x n row-major; BT: op(B) = B^T.  64 x 64 tile per CTA, 4 x 4 per thread.
template <bool BT>
__global__ void __launch_bounds__(256)
k_bgemm(const double* __restrict__ A, int64_t strideA, const double* __restrict__ Bm, int64_t strideB,
        const int32_t* __restrict__ recB, const double* __restrict__ sc, double* __restrict__ C, int64_t strideC, int n) {
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const int b = blockIdx.z;
  const double* Ab = A + b * strideA;
  const double* Bb = Bm + (recB ? (int64_t)recB[b] : (int64_t)b) * strideB;
  if (recB && recB[b] < 0) return;
  const int i0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
  for (int m0 = 0; m0 < n; m0 += 16) {
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += 256) {
      int ii = e / 16, mm = e % 16;
      double v = 0.0;
      if (i0 + ii < n && m0 + mm < n) {
        v = Ab[(int64_t)(i0 + ii) * n + m0 + mm];
        if (sc) v *= sc[(int64_t)b * n + m0 + mm];
      }
      As[mm][ii] = v;
      double w = 0.0;
      if (BT) {
        if (k0 + ii < n && m0 + mm < n) w = Bb[(int64_t)(k0 + ii) * n + m0 + mm];
        Bs[mm][ii] = w;
      } else {
        int mm2 = e / 64, kk = e % 64;
        if (m0 + mm2 < n && k0 + kk < n) w = Bb[(int64_t)(m0 + mm2) * n + k0 + kk];
        Bs[mm2][kk] = w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = As[mm][ty + 16 * a];
#pragma unroll
      for (int c = 0; c < 4; ++c) bf[c] = Bs[mm][tx + 16 * c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fma(af[a], bf[c], acc[a][c]);
    }
  }
  double* Cb = C + b * strideC;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int i = i0 + ty + 16 * a;
    if (i >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int k = k0 + tx + 16 * c;
      if (k < n) Cb[(int64_t)i * n + k] = acc[a][c];
    }
  }
}

__global__ void k_cov_nan(int64_t s0, int n, SysBuf B, double* __restrict__ dC_chunk) {
  const int64_t s = s0 + blockIdx.x;
  if (B.st[s] == VI_ST_OK && B.rec[s] >= 0) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double* out = dC_chunk + (int64_t)blockIdx.x * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) out[e] = nan;
}

// Fallback for orders whose vectors do not fit shared memory: everything in one thread per system
// on the global (warp-interleaved) workspace.
__global__ void __launch_bounds__(64)
k_tql(int64_t nsys, SysBuf B, double rcond, double* __restrict__ Cout, int32_t* __restrict__ rank_out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsys) return;
  const int st = B.st[s];
  if (st == kSkip) return;
  const int n = B.n;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double* Cs = Cout + s * (int64_t)n;
  if (st != VI_ST_OK) {
    for (int i = 0; i < n; ++i) Cs[i] = nan;
    rank_out[s] = 0;
    return;
  }
  const int64_t base = ileave(s, n);
  vi_svec d{B.d + base, 32}, e{B.e + base, 32}, g{B.g + base, 32};
  vi_tape tape = tape_of(B, s);
  int32_t nrot = 0;
  const int q = vi_tql(n, d, e, g, tape, &nrot);
  if (q != 0) {
    B.st[s] = VI_ST_NOCONV;
    for (int i = 0; i < n; ++i) Cs[i] = nan;
    rank_out[s] = 0;
    return;
  }
  const int rank = vi_spectral_divide(n, d, g, rcond);
  vi_tape_apply_z(g, tape, nrot);
  const double scl = B.scl[s];
  for (int i = 0; i < n; ++i) g[i] = g[i] * scl;
  vi_tri_backtransform(n, B.V + s * B.vstride, B.tau + base, 32, g.p, g.stride);
  for (int i = 0; i < n; ++i) Cs[i] = g[i];
  rank_out[s] = rank;
}

// chi2[s] = sum_j Wm[r][j] * (sum_n At[n][j] C_s[n] - bm[r][j])^2   (chi2objfunct, interpolate.py:258-259)
// grid (systems/16, gate chunks): a CTA contracts kChiGates gates against 16 coefficient vectors (the
// design matrix is streamed once per 16 systems) and writes one partial sum per system; k_chi2_sum
// adds the partials in chunk order (deterministic).
__global__ void __launch_bounds__(kChiThreads)
k_chi2(const double* __restrict__ At, const double* __restrict__ Wm, const double* __restrict__ bm, int P, int n,
       int64_t nsys, const int32_t* __restrict__ rec, const int32_t* __restrict__ st, const double* __restrict__ Csys,
       double* __restrict__ part, int nchunk) {
  extern __shared__ __align__(16) double sm[];
  double* Cs = sm;                         // n x kChiSB  (n-major so one LDS.128 serves two systems)
  double* red = sm + (size_t)n * kChiSB;   // kChiThreads/32 x kChiSB
  __shared__ int srec[kChiSB];
  const int64_t s0 = (int64_t)blockIdx.x * kChiSB;
  const int chunk = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < kChiSB) {
    int64_t s = s0 + tid;
    int r = kSkip;
    if (s < nsys && st[s] == VI_ST_OK) r = rec[s];
    srec[tid] = r;
  }
  __syncthreads();
  bool any = false;
  for (int q = 0; q < kChiSB; ++q) any = any || (srec[q] >= 0);
  if (!any) return;
  for (int e = tid; e < n * kChiSB; e += kChiThreads) {
    int i = e / kChiSB, q = e - i * kChiSB;
    Cs[e] = (srec[q] >= 0) ? Csys[(s0 + q) * (int64_t)n + i] : 0.0;
  }
  __syncthreads();
  bool same = true;
  for (int q = 1; q < kChiSB; ++q) same = same && (srec[q] == srec[0]);
  double csum[kChiSB];
#pragma unroll
  for (int q = 0; q < kChiSB; ++q) csum[q] = 0.0;
  const int jend = min(P, (chunk + 1) * kChiGates);
  for (int j = chunk * kChiGates + tid; j < jend; j += kChiThreads) {
    double acc[kChiSB];
#pragma unroll
    for (int q = 0; q < kChiSB; ++q) acc[q] = 0.0;
    const double* ap = At + j;
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const double a = __ldg(ap + (int64_t)i * P);
      const double2* c2 = reinterpret_cast<const double2*>(Cs + i * kChiSB);
#pragma unroll
      for (int q = 0; q < kChiSB / 2; ++q) {
        double2 c = c2[q];
        acc[2 * q] = fma(a, c.x, acc[2 * q]);
        acc[2 * q + 1] = fma(a, c.y, acc[2 * q + 1]);
      }
    }
    if (same) {
      const int r = srec[0];
      const double w = Wm[(int64_t)r * P + j], b = bm[(int64_t)r * P + j];
#pragma unroll
      for (int q = 0; q < kChiSB; ++q) { double t = acc[q] - b; csum[q] += (t * t) * w; }
    } else {
#pragma unroll
      for (int q = 0; q < kChiSB; ++q) {
        const int r = srec[q];
        if (r >= 0) {
          const double w = Wm[(int64_t)r * P + j], b = bm[(int64_t)r * P + j];
          double t = acc[q] - b;
          csum[q] += (t * t) * w;
        }
      }
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int q = 0; q < kChiSB; ++q) {
    double v = csum[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * kChiSB + q] = v;
  }
  __syncthreads();
  if (tid < kChiSB) {
    int64_t s = s0 + tid;
    if (s < nsys && srec[tid] >= 0) {
      double v = 0.0;
      for (int w = 0; w < kChiThreads / 32; ++w) v += red[w * kChiSB + tid];
      part[s * nchunk + chunk] = v;
    }
  }
}

__global__ void k_chi2_sum(int64_t nsys, const int32_t* __restrict__ st, const double* __restrict__ part, int nchunk,
                           double* __restrict__ chi2) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsys || st[s] != VI_ST_OK) return;
  double v = 0.0;
  for (int c = 0; c < nchunk; ++c) v += part[s * nchunk + c];
  chi2[s] = v;
}

// ---- system set-up for the three phases ---------------------------------------------------
// Table phase, LAZY: the decade walk (interpolate.py:180-207) reads chi2(10^-k) for k = 0, 1, ... and stops at
// the first sign change of chi2 - nu, so only the entries up to there are ever needed.  The table is extended in
// passes of `step` decades per unit; after each pass the walk (first scale factor) is replayed on what exists
// and the unit leaves the table phase as soon as it would not read further.  A unit whose walk runs off the
// end gets its full table (the later scale factors re-read it).  Decisions are identical to evaluating all
// 102 entries up front; only entries the reference never looks at are skipped.
__device__ __forceinline__ int table_limit(int ks) { return (ks < VI_NALPHA - 1 ? ks : VI_NALPHA - 1) + 1; }

__global__ void k_table_init(int64_t U, int nreg, const int32_t* __restrict__ npts, UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  Ub.kdone[u] = 0;
  Ub.walking[u] = npts[u / nreg] > 0 ? 1 : 0;
}

__global__ void k_table_plan(int64_t U, int step, UnitBuf Ub, int64_t* __restrict__ cnt) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int c = 0;
  if (Ub.walking[u]) {
    c = table_limit(Ub.kstar[u]) - Ub.kdone[u];
    if (c > step) c = step;
  }
  cnt[u] = c;
}

// after a pass: account for the new entries, replicate entries k > kstar (bit-identical systems) once the
// distinct ones are complete, and replay the walk to see whether it needs more
__global__ void k_table_advance(int64_t U, int nreg, const int32_t* __restrict__ npts, const int64_t* __restrict__ cnt,
                                UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U || !Ub.walking[u]) return;
  const int kd = Ub.kdone[u] + (int)cnt[u];
  Ub.kdone[u] = kd;
  const int lim = table_limit(Ub.kstar[u]);
  double* tab = Ub.table + u * VI_NALPHA;
  int avail = kd;
  if (kd >= lim) {
    const double v = tab[lim - 1];
    for (int k = lim; k < VI_NALPHA; ++k) tab[k] = v;
    avail = VI_NALPHA;
  }
  Ub.walking[u] = vi_chi2_walk_needs_more(tab, 1, npts[u / nreg], avail) ? 1 : 0;
}

// table phase, compact numbering: system t in [0, off[U]) belongs to unit u = last unit with off[u] <= t,
// k = kdone[u] + t - off[u]; the chunk covers [t0, t0 + cnt)
__global__ void k_setup_table(int64_t t0, int64_t cnt, int64_t U, int nreg, const double* __restrict__ pow10tab,
                              const int64_t* __restrict__ off, const int32_t* __restrict__ kdone, SysBuf B) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.cap) return;
  if (s >= cnt) { B.rec[s] = kSkip; return; }
  const int64_t t = t0 + s;
  int64_t lo = 0, hi = U;              // off[lo] <= t < off[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= t) lo = mid; else hi = mid;
  }
  const int64_t u = lo;
  const int k = kdone[u] + (int)(t - off[u]);
  const int r = (int)(u / nreg), q = (int)(u - (int64_t)r * nreg);
  B.rec[s] = r;
  B.unit[s] = (int32_t)u;
  B.kidx[s] = k;
  for (int i = 0; i < nreg; ++i) B.lam[s * nreg + i] = (i == q) ? pow10tab[k] : 0.0;
}

// kstar[u] = smallest k for which fl(sym(G) + 10^-k R) == sym(G) in every entry (VI_NALPHA if none).
// Rounding is monotone, so every k >= kstar gives the bit-identical system; the reference solves all
// of them again and gets the same chi^2 (interpolate.py:193-203 walks alpha down to -101).
__global__ void __launch_bounds__(256)
k_kstar(int nreg, int n, const double* __restrict__ G, const double* __restrict__ regs,
        const double* __restrict__ pow10tab, int32_t* __restrict__ kstar) {
  __shared__ int sflag;
  const int64_t u = blockIdx.x;
  const int r = (int)(u / nreg), q = (int)(u - (int64_t)r * nreg);
  const double* Gr = G + (int64_t)r * n * n;
  const double* Rq = regs + (int64_t)q * n * n;
  int lo = 0, hi = VI_NALPHA;       // invariant: systems with k >= hi are identical to G (hi = NALPHA: none known)
  // bisection on the monotone predicate same(k)
  while (lo < hi) {
    const int k = (lo + hi) / 2;
    const double l = pow10tab[k];
    if (threadIdx.x == 0) sflag = 1;
    __syncthreads();
    int same = 1;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      int i = idx / n, c = idx - i * n;
      double x = 0.5 * (Gr[(int64_t)i * n + c] + Gr[(int64_t)c * n + i]);
      double x2 = fma(l, Rq[idx], x);      // same expression as vi_tri_load
      if (x2 != x) same = 0;
    }
    if (!same) sflag = 0;
    __syncthreads();
    const int all = sflag;
    __syncthreads();
    if (all) hi = k; else lo = k + 1;
  }
  if (threadIdx.x == 0) kstar[u] = hi;
}

__global__ void k_scatter_table(int64_t cnt, SysBuf B, UnitBuf Ub) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cnt) return;
  int st = B.st[s];
  if (st == kSkip) return;
  const int64_t u = B.unit[s];
  const int64_t t = u * VI_NALPHA + B.kidx[s];
  if (st != VI_ST_OK) { atomicMax(&Ub.tabbad[u], st); Ub.table[t] = __longlong_as_double(0x7ff8000000000000LL); return; }
  Ub.table[t] = B.chi2[s];
}

__global__ void k_bracket(int64_t U, int nreg, const int32_t* __restrict__ npts, UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int r = (int)(u / nreg);
  Ub.active[u] = 0;
  Ub.nu[u] = 0.0;
  Ub.klo[u] = -1;
  if (npts[r] <= 0) { Ub.status[u] = VI_ST_EMPTY; return; }
  const double* tab = Ub.table + u * VI_NALPHA;
  vi_bracket br = vi_chi2_bracket(tab, 1, npts[r]);
  if (Ub.tabbad[u] != 0) {
    // a table system failed: it only matters if the walk actually read that entry (stored as NaN)
    const int last = (br.status == VI_ST_OK) ? br.k_lo : (br.status == VI_ST_TOO_SMOOTH ? 0 : VI_NALPHA - 1);
    for (int k = 0; k <= last; ++k)
      if (isnan(tab[k])) { Ub.status[u] = Ub.tabbad[u]; return; }
  }
  Ub.status[u] = br.status;
  Ub.nu[u] = br.nu;
  Ub.klo[u] = br.k_lo;
  if (br.status != VI_ST_OK) return;
  const int k = br.k_lo;
  vi_brent b;
  vi_brent_init(b, -(double)k, tab[k] - br.nu, -(double)(k - 1), tab[k - 1] - br.nu);
  Ub.br[u] = b;
  Ub.active[u] = 1;
}

// Brent: advance units [u0, u0+cnt) to their next abscissa; units still searching get a compact system
// slot (atomic counter: slot order is arbitrary, results do not depend on it)
__global__ void k_brent_propose(int64_t u0, int64_t cnt, int nreg, SysBuf B, UnitBuf Ub) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  int64_t u = u0 + i;
  if (!Ub.active[u]) return;
  vi_brent b = Ub.br[u];
  bool fin = vi_brent_propose(b);
  Ub.br[u] = b;
  if (fin) {
    Ub.active[u] = 0;
    if (b.done != 1) Ub.status[u] = VI_ST_NOCONV;
    return;
  }
  int r = (int)(u / nreg), q = (int)(u - (int64_t)r * nreg);
  const int64_t s = atomicAdd(Ub.count, 1);
  B.rec[s] = r;
  B.unit[s] = (int32_t)u;
  const double lamv = exp10(b.xcur);   // np.power(10., alpha), interpolate.py:250
  for (int k = 0; k < nreg; ++k) B.lam[s * nreg + k] = (k == q) ? lamv : 0.0;
}

__global__ void k_brent_feed(int64_t cnt, SysBuf B, UnitBuf Ub) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cnt) return;
  int64_t u = B.unit[s];
  int st = B.st[s];
  if (st != VI_ST_OK) { Ub.active[u] = 0; Ub.status[u] = (st == VI_ST_NONFINITE) ? VI_ST_NONFINITE : VI_ST_NOCONV; return; }
  double f = B.chi2[s] - Ub.nu[u];
  if (!isfinite(f)) { Ub.active[u] = 0; Ub.status[u] = VI_ST_NOCONV; return; }
  vi_brent b = Ub.br[u];
  vi_brent_feed(b, f);
  Ub.br[u] = b;
}

// ---- GCV (interpolate.py:263-351): leave-one-gate-out residual sum minimised by Nelder-Mead -------
// validx[r][k] = index of the k-th gate of record r that carries weight (a zero-weight gate adds nothing
// to the objective and leaves the system unchanged, so it is skipped)
__global__ void __launch_bounds__(32)
k_valid_index(int P, const double* __restrict__ Wm, int32_t* __restrict__ validx, int32_t* __restrict__ nvalid) {
  const int r = blockIdx.x, lane = threadIdx.x;
  int base = 0;
  for (int j0 = 0; j0 < P; j0 += 32) {
    const int j = j0 + lane;
    const bool ok = (j < P) && (Wm[(int64_t)r * P + j] != 0.0);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) validx[(int64_t)r * P + base + __popc(m & ((1u << lane) - 1u))] = j;
    base += __popc(m);
  }
  if (lane == 0) nvalid[r] = base;
}

__global__ void k_nm_init(int64_t U, int nreg, const int32_t* __restrict__ npts, UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  vi_nm s;
  vi_nm_init(s, -20.0);                      // alpha0, interpolate.py:288
  Ub.nm[u] = s;
  const bool ok = npts[u / nreg] > 0;
  Ub.active[u] = ok ? 1 : 0;
  Ub.status[u] = ok ? VI_ST_OK : VI_ST_EMPTY;
}

// one objective evaluation per active unit: next abscissa, number of systems it needs
__global__ void k_nm_propose(int64_t U, int nreg, const int32_t* __restrict__ nvalid, UnitBuf Ub, int64_t* __restrict__ cnt) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  cnt[u] = 0;
  if (!Ub.active[u]) return;
  vi_nm s = Ub.nm[u];
  double x = 0.0;
  const bool more = vi_nm_next(s, &x);
  Ub.nm[u] = s;
  if (!more) {
    Ub.active[u] = 0;
    Ub.status[u] = s.success ? VI_ST_OK : VI_ST_NOCONV;   // 'Minima of GCV function could not be found' -> NaN
    Ub.br[u].root = s.x0;
    return;
  }
  Ub.alpha[u] = x;
  Ub.fsum[u] = 0.0;
  cnt[u] = nvalid[u / nreg];
}

__global__ void __launch_bounds__(1024)
k_scan_counts(int64_t U, const int64_t* __restrict__ cnt, int64_t* __restrict__ off) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int64_t per = (U + 1023) / 1024;
  const int64_t a = t * per, b = (a + per < U) ? a + per : U;
  int64_t sum = 0;
  for (int64_t u = a; u < b; ++u) sum += cnt[u];
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    int64_t run = 0;
    for (int i = 0; i < 1024; ++i) { int64_t v = part[i]; part[i] = run; run += v; }
    off[U] = run;
  }
  __syncthreads();
  int64_t run = part[t];
  for (int64_t u = a; u < b; ++u) { off[u] = run; run += cnt[u]; }
}

__global__ void k_gcv_setup(int64_t t0, int64_t cnt, int64_t U, int nreg, int P, const int64_t* __restrict__ off,
                            const int32_t* __restrict__ validx, SysBuf B, UnitBuf Ub) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.cap) return;
  if (s >= cnt) { B.rec[s] = kSkip; return; }
  const int64_t t = t0 + s;
  int64_t lo = 0, hi = U;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= t) lo = mid; else hi = mid;
  }
  const int64_t u = lo;
  const int k = (int)(t - off[u]);
  const int r = (int)(u / nreg), q = (int)(u - (int64_t)r * nreg);
  B.rec[s] = r;
  B.unit[s] = (int32_t)u;
  B.kidx[s] = k;
  B.gate[s] = validx[(int64_t)r * P + k];
  const double lamv = exp10(Ub.alpha[u]);
  for (int i = 0; i < nreg; ++i) B.lam[s * nreg + i] = (i == q) ? lamv : 0.0;
}

// residual of the left-out gate: (a_j . C - b_j)^2 w_j   (interpolate.py:347-349), one warp per system
__global__ void __launch_bounds__(256)
k_gcv_resid(int64_t cnt, int n, Downdate dd, SysBuf B, const double* __restrict__ Csys, double* __restrict__ res) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= cnt) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (B.st[s] != VI_ST_OK) { if (lane == 0) res[s] = nan; return; }
  const int r = B.rec[s], j = B.gate[s];
  const double* a = dd.A + (int64_t)j * n;
  const double* c = Csys + s * (int64_t)n;
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) acc += a[i] * c[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double d = acc - dd.bm[(int64_t)r * dd.P + j];
    res[s] = (d * d) * dd.Wm[(int64_t)r * dd.P + j];
  }
}

// objective[u] += residuals of the unit's systems inside this chunk, in gate order (the reference sums
// them with Python's sequential sum, interpolate.py:351)
__global__ void k_gcv_accumulate(int64_t U, int64_t t0, int64_t cnt, const int64_t* __restrict__ off,
                                 const double* __restrict__ res, UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int64_t a = off[u], b = off[u + 1];
  if (a < t0) a = t0;
  if (b > t0 + cnt) b = t0 + cnt;
  if (a >= b) return;
  double acc = Ub.fsum[u];
  for (int64_t t = a; t < b; ++t) acc += res[t - t0];
  Ub.fsum[u] = acc;
}

__global__ void k_nm_feed(int64_t U, UnitBuf Ub) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U || !Ub.active[u]) return;
  vi_nm s = Ub.nm[u];
  vi_nm_feed(s, Ub.fsum[u]);
  Ub.nm[u] = s;
}

// final phase: records [r0, r0+cnt), slot s = r - r0
__global__ void k_setup_final(int64_t r0, int64_t cnt, int nreg, int method, const int32_t* __restrict__ npts,
                              SysBuf B, UnitBuf Ub, double* __restrict__ lam_out, int32_t* __restrict__ status_out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.cap) return;
  if (s >= cnt) { B.rec[s] = kSkip; return; }
  int64_t r = r0 + s;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  int worst = VI_ST_OK;
  bool bad = false;
  if (npts[r] <= 0) { worst = VI_ST_EMPTY; bad = true; }
  if (method != VI_METHOD_NONE) {
    for (int q = 0; q < nreg; ++q) {
      int64_t u = r * nreg + q;
      int st = Ub.status[u];
      double l;
      if (bad) l = nan;
      else if (st == VI_ST_OK) l = exp10(Ub.br[u].root);   // interpolate.py:216
      else if (st == VI_ST_TOO_SMOOTH) l = 0.0;
      else l = nan;
      if (!bad && st != VI_ST_OK && st != VI_ST_TOO_SMOOTH) { bad = true; worst = st; }
      else if (!bad && st == VI_ST_TOO_SMOOTH && worst == VI_ST_OK) worst = VI_ST_TOO_SMOOTH;
      lam_out[r * nreg + q] = l;
    }
  } else {
    for (int q = 0; q < nreg; ++q) lam_out[r * nreg + q] = 0.0;
  }
  status_out[r] = worst;
  B.rec[s] = bad ? kSkip : (int)r;
  for (int q = 0; q < nreg; ++q) B.lam[s * nreg + q] = bad ? 0.0 : lam_out[r * nreg + q];
}

__global__ void k_finalize(int64_t r0, int64_t cnt, int n, SysBuf B, double* __restrict__ C, double* __restrict__ chi2,
                           int32_t* __restrict__ rank, int32_t* __restrict__ status_out) {
  int64_t s = (int64_t)blockIdx.x;
  if (s >= cnt) return;
  int64_t r = r0 + s;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  int st = B.st[s];
  bool bad = (B.rec[s] < 0) || (st != VI_ST_OK);
  if (bad) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) C[r * n + i] = nan;
    if (threadIdx.x == 0) {
      chi2[r] = nan;
      rank[r] = 0;
      if (B.rec[s] >= 0 && st != VI_ST_OK) status_out[r] = (st == VI_ST_NONFINITE) ? VI_ST_NONFINITE : VI_ST_NOCONV;
    }
  } else if (threadIdx.x == 0) {
    chi2[r] = B.chi2[s];
    rank[r] = B.rank[s];
  }
}

__global__ void k_setup_solve(int64_t s0, int64_t cnt, int nreg, const int32_t* __restrict__ rec,
                              const double* __restrict__ lam, SysBuf B) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.cap) return;
  if (s >= cnt) { B.rec[s] = kSkip; return; }
  B.rec[s] = rec ? rec[s0 + s] : (int)(s0 + s);
  for (int q = 0; q < nreg; ++q) B.lam[s * nreg + q] = lam[(s0 + s) * nreg + q];
}

__global__ void k_copy_status(int64_t cnt, SysBuf B, int32_t* __restrict__ rank, int32_t* __restrict__ status) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cnt) return;
  int st = B.st[s];
  status[s] = (st == kSkip) ? VI_ST_EMPTY : st;
  rank[s] = (st == VI_ST_OK) ? B.rank[s] : 0;
}

__global__ void k_transpose(const double* __restrict__ A, int P, int N, double* __restrict__ At) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)P * N) return;
  int j = (int)(e / N), i = (int)(e - (int64_t)j * N);
  At[(int64_t)i * P + j] = A[e];
}

// ------------------------------------------------------------------------------------------
// host-side drivers
// ------------------------------------------------------------------------------------------
inline unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

int run_tridiag(int64_t cnt, const double* G, const double* y, const double* regs, const SysBuf& B, cudaStream_t s,
                Downdate dd = Downdate{nullptr, nullptr, nullptr, 0}) {
  if (cnt <= 0) return VI_OK;
  if (B.two_stage && !two_stage_smem(B.n)) {
    const size_t smem1 = (size_t)vi_bnd_doubles_big(B.n, kBigWarps) * sizeof(double);
    VI_CUDA(cudaFuncSetAttribute(k_band_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    VI_KERNEL(VI_K_TRIDIAG, s, k_band_big<<<(unsigned)cnt, kBigWarps * 32, smem1, s>>>(G, y, regs, B, dd));
    const size_t per = (size_t)vi_chs_doubles(B.n) * sizeof(double);
    int cw = (int)((227 * 1024) / per);
    if (cw > kChaseWarps) cw = kChaseWarps;
    if (cw > 1 && (227 * 1024) / (per * cw) < 2 && (227 * 1024) / per >= 2) cw = 1;      // rather two CTAs of one warp than one of two
    const size_t smem2 = per * cw;
    VI_CUDA(cudaFuncSetAttribute(k_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    VI_KERNEL(VI_K_CHASE, s, k_chase<<<blocks(cnt, cw), cw * 32, smem2, s>>>(cnt, B));
    return VI_OK;
  }
  if (B.two_stage) {
    const size_t smem1 = (size_t)vi_bnd_doubles(B.n) * sizeof(double);
    const int nt = vi_bnd_threads(B.n);
    if (nt == 32 * VI_BND_NW) {
      VI_CUDA(cudaFuncSetAttribute(k_band<32 * VI_BND_NW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      VI_KERNEL(VI_K_TRIDIAG, s, (k_band<32 * VI_BND_NW, 2><<<(unsigned)cnt, nt, smem1, s>>>(G, y, regs, B, dd)));
    } else {
      VI_CUDA(cudaFuncSetAttribute(k_band<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      VI_KERNEL(VI_K_TRIDIAG, s, (k_band<128, 1><<<(unsigned)cnt, nt, smem1, s>>>(G, y, regs, B, dd)));
    }
    if (B.split_p1 > 0) {
      const size_t smemt = (size_t)vi_bnd_doubles(B.n - 8 * B.split_p1, kTailWarps) * sizeof(double);
      VI_CUDA(cudaFuncSetAttribute(k_band_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemt));
      VI_KERNEL(VI_K_TRIDIAG, s, k_band_tail<<<(unsigned)cnt, kTailWarps * 32, smemt, s>>>(B));
    }
    const size_t smem2 = (size_t)kChaseWarps * vi_chs_doubles(B.n) * sizeof(double);
    VI_CUDA(cudaFuncSetAttribute(k_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    VI_KERNEL(VI_K_CHASE, s, k_chase<<<blocks(cnt, kChaseWarps), kChaseWarps * 32, smem2, s>>>(cnt, B));
    return VI_OK;
  }
  {
    // packed form whenever its working set fits one CTA (two CTAs per SM up to n = 144)
    static const bool force_full = getenv("VI_TRIDIAG_FULL") != nullptr;
    const size_t smem = (size_t)vi_trp_doubles(B.n) * sizeof(double);
    const int nt = vi_trp_threads(B.n);
    if (!force_full && smem <= 227 * 1024 && nt <= 448) {
      if (nt <= 320) {       // n <= 144: two CTAs per SM (registers capped accordingly)
        VI_CUDA(cudaFuncSetAttribute(k_tridiag_packed<320, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VI_KERNEL(VI_K_TRIDIAG, s, (k_tridiag_packed<320, 2><<<(unsigned)cnt, nt, smem, s>>>(G, y, regs, B, dd)));
      } else {
        VI_CUDA(cudaFuncSetAttribute(k_tridiag_packed<448, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VI_KERNEL(VI_K_TRIDIAG, s, (k_tridiag_packed<448, 1><<<(unsigned)cnt, nt, smem, s>>>(G, y, regs, B, dd)));
      }
      return VI_OK;
    }
  }
  if (B.use_gx) {
    VI_CUDA(cudaFuncSetAttribute(k_tridiag<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
    VI_KERNEL(VI_K_TRIDIAG, s, k_tridiag<true><<<(unsigned)cnt, B.nt, B.smem, s>>>(G, y, regs, B, dd));
  } else {
    VI_CUDA(cudaFuncSetAttribute(k_tridiag<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
    VI_KERNEL(VI_K_TRIDIAG, s, k_tridiag<false><<<(unsigned)cnt, B.nt, B.smem, s>>>(G, y, regs, B, dd));
  }
  return VI_OK;
}

// QL (thread per system) on stream s_ql, apply (warp per system) on stream s_apply.  The QL kernel takes
// the whole shared memory of an SM, so it serialises with k_tridiag; k_apply and k_chi2 are small and can
// run next to the tridiagonalisation of the following chunk (pipelined table phase: s_apply != s_ql, the
// caller orders the two streams with events).
// lanes per warp that carry a system: as few as keep every system of the batch resident at once, at most Lmax
int sparse_lanes(int64_t cnt, int sys_per_sm_max, int warps_per_sm, int Lmax) {
  int64_t per_sm = (cnt + sm_count() - 1) / sm_count();
  if (per_sm > sys_per_sm_max) per_sm = sys_per_sm_max;
  int L = (int)((per_sm + warps_per_sm - 1) / warps_per_sm);
  if (L < 1) L = 1;
  if (L > Lmax) L = Lmax;
  return L;
}

int run_ql(int64_t cnt, const SysBuf& B, cudaStream_t s, bool* split) {
  const size_t per_sys = (size_t)2 * B.n * sizeof(double);
  const int W = ql_warps(B.n), Lmax = ql_lanes_max(B.n);
  *split = W > 0;
  if (cnt <= 0 || !*split) return VI_OK;
  const int L = sparse_lanes(cnt, W * Lmax, W, Lmax);
  const int T = W * L;
  size_t smem = (size_t)T * per_sys;
  VI_CUDA(cudaFuncSetAttribute(k_tql_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // lanes that reach the (expensive) sweep set-up wait until this many of the warp are there, or nobody rotates
  static const int batch_env = env_int("VI_TQL_BATCH", 0);
  const int iter_batch = batch_env > 0 ? batch_env : (L >= 32 ? 8 : 1);     // sparse lanes: waiting does not pay (measured)
  // few systems per SM: one system per warp (k_tql_single), as many warps as there are systems.  Past ~16-24 warps per
  // SM the single-lane form becomes issue-bound (a warp instruction serves one system) and the sparse-lane state
  // machine wins again.
  static const int single_max = env_int("VI_TQL_SINGLE_MAX", 16);
  const int64_t per_sm = (cnt + sm_count() - 1) / sm_count();
  if (per_sm <= single_max) {
    int W1 = (int)per_sm;
    if (W1 > 24) W1 = 24;      // 78 registers per thread
    if (W1 > ql_fit(B.n)) W1 = ql_fit(B.n);
    if (W1 < 1) W1 = 1;
    const size_t smem1 = (size_t)W1 * per_sys;
    VI_CUDA(cudaFuncSetAttribute(k_tql_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    VI_KERNEL(VI_K_TQL, s, k_tql_single<<<blocks(cnt, W1), W1 * 32, smem1, s>>>(cnt, B));
    return VI_OK;
  }
  VI_KERNEL(VI_K_TQL, s, k_tql_smem<<<blocks(cnt, T), W * 32, smem, s>>>(cnt, B, L, iter_batch));
  return VI_OK;
}

int run_apply(int64_t cnt, const SysBuf& B, double rcond, double* Cout, int32_t* rank_out, cudaStream_t s, bool split) {
  if (cnt <= 0) return VI_OK;
  static const bool old_replay = getenv("VI_OLD_REPLAY") != nullptr;
  // four systems per warp: measured SLOWER than one (10.8 vs 8.7 ms per 28 416 systems: with 8 lanes per system the
  // wavefront needs 3x the steps on these spectra); kept behind VI_WAVE4_MIN=<batch size> for other workloads
  static const int wave4_min = env_int("VI_WAVE4_MIN", 0);
  if (split && !old_replay && B.two_stage && B.wtab && wave4_min > 0 && cnt >= wave4_min) {
    const size_t smem = (size_t)kWaveWarps * kWaveSys * (((B.n + 7) & ~7) + 8) * sizeof(double);
    if (smem <= 227 * 1024) {
      VI_CUDA(cudaFuncSetAttribute(k_replay_wave4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VI_KERNEL(VI_K_APPLY, s, k_replay_wave4<<<blocks(cnt, kWaveWarps * kWaveSys), kWaveWarps * 32, smem, s>>>(cnt, B, rcond, Cout, rank_out));
      return VI_OK;
    }
  }
  if (split && !old_replay) {
    const size_t smem = (size_t)kWaveWarps * (B.wtab ? vi_wav_bytes_vec(B.n) : vi_wav_bytes(B.n));
    VI_CUDA(cudaFuncSetAttribute(k_replay_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VI_KERNEL(VI_K_APPLY, s, k_replay_wave<<<blocks(cnt, kWaveWarps), kWaveWarps * 32, smem, s>>>(cnt, B, rcond, Cout, rank_out));
    return VI_OK;
  }
  if (split) {
    const size_t per_sys = (size_t)B.n * sizeof(double);
    const int fit = (int)((227 * 1024) / per_sys);
    static const int Lenv = env_int("VI_REPLAY_LANES", 16);
    static const int Wenv = env_int("VI_REPLAY_WARPS", 8);
    const int kReplayWarps = Wenv > kReplayWarpsMax ? kReplayWarpsMax : Wenv;
    int Lmax = Lenv > 32 ? 32 : Lenv;
    while (Lmax > 1 && (size_t)kReplayWarps * Lmax * per_sys > 227 * 1024) --Lmax;
    int nb = 1;      // resident CTAs per SM at the full lane count (registers and shared memory both count)
    VI_CUDA(cudaFuncSetAttribute(k_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kReplayWarps * Lmax * per_sys)));
    VI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_replay, kReplayWarps * 32, (size_t)kReplayWarps * Lmax * per_sys));
    if (nb < 1) nb = 1;
    const int L = sparse_lanes(cnt, nb * kReplayWarps * Lmax, nb * kReplayWarps, Lmax);
    (void)fit;
    size_t smem1 = (size_t)kReplayWarps * L * per_sys;
    VI_CUDA(cudaFuncSetAttribute(k_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    VI_KERNEL(VI_K_APPLY, s, k_replay<<<blocks(cnt, kReplayWarps * L), kReplayWarps * 32, smem1, s>>>(cnt, B, rcond, L));
    if (B.two_stage) {
      size_t smem2 = (size_t)kChaseWarps * (vi_bnd_npad(B.n) + 8) * sizeof(double);
      VI_KERNEL(VI_K_APPLY, s, k_apply_ts<<<blocks(cnt, kChaseWarps), kChaseWarps * 32, smem2, s>>>(cnt, B, Cout, rank_out));
    } else {
      size_t smem2 = (size_t)kApplyWarps * B.n * sizeof(double);
      VI_KERNEL(VI_K_APPLY, s, k_apply<<<blocks(cnt, kApplyWarps), kApplyWarps * 32, smem2, s>>>(cnt, B, Cout, rank_out));
    }
  } else {
    VI_KERNEL(VI_K_TQL, s, k_tql<<<blocks(cnt, 64), 64, 0, s>>>(cnt, B, rcond, Cout, rank_out));
  }
  return VI_OK;
}

int run_post(int64_t cnt, const SysBuf& B, double rcond, double* Cout, int32_t* rank_out, cudaStream_t s) {
  bool split = false;
  if (int rc = run_ql(cnt, B, s, &split)) return rc;
  return run_apply(cnt, B, rcond, Cout, rank_out, s, split);
}

int run_systems(int64_t cnt, const double* G, const double* y, const double* regs, const SysBuf& B, double rcond,
                double* Cout, int32_t* rank_out, cudaStream_t s) {
  if (int rc = run_tridiag(cnt, G, y, regs, B, s)) return rc;
  return run_post(cnt, B, rcond, Cout, rank_out, s);
}

int run_chi2(int64_t cnt, const double* At, const double* Wm, const double* bm, int P, const SysBuf& B,
             const double* Csys, double* chi2, cudaStream_t s) {
  if (cnt <= 0) return VI_OK;
  size_t smem = ((size_t)B.n * kChiSB + (kChiThreads / 32) * kChiSB) * sizeof(double);
  VI_CUDA(cudaFuncSetAttribute(k_chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(blocks(cnt, kChiSB), (unsigned)B.nchunk);
  VI_KERNEL(VI_K_CHI2, s, k_chi2<<<grid, kChiThreads, smem, s>>>(At, Wm, bm, P, B.n, cnt, B.rec, B.st, Csys, B.chi2p, B.nchunk));
  VI_KERNEL(VI_K_CHI2, s, k_chi2_sum<<<blocks(cnt, 256), 256, 0, s>>>(cnt, B.st, B.chi2p, B.nchunk, chi2));
  return VI_OK;
}

// dC chunk = H (A^T W A) H for the systems [c0, c0 + nc) of the current batch (interpolate.py:464-467): eigenvectors,
// H = E diag(scl / lambda | cut-off N eps max|lambda|) E^T, T = H G, out = T H; NaN blocks for failed systems.
int run_cov_chunk(int64_t c0, int64_t nc, int N, const SysBuf& B, const double* G, double* covE, double* covH,
                  double* covT, double* covD, double* out, cudaStream_t st) {
  const int64_t NN = (int64_t)N * N;
  size_t smem_e = ((size_t)N * B.ld + eigvec_stage_doubles(N)) * sizeof(double);
  const bool ev_global = smem_e > 227 * 1024;       // N > ~160: eigenvector block stays in global memory
  if (ev_global) smem_e = (size_t)eigvec_stage_doubles(N) * sizeof(double);
  if (ev_global) VI_CUDA(cudaFuncSetAttribute(k_eigvec<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
  else VI_CUDA(cudaFuncSetAttribute(k_eigvec<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
  const unsigned tiles = (unsigned)((N + 63) / 64);
  VI_CUDA(cudaMemsetAsync(covD, 0, nc * N * sizeof(double), st));
  if (ev_global)
    VI_KERNEL(VI_K_COV, st, k_eigvec<true><<<(unsigned)nc, (N + 31) / 32 * 32, smem_e, st>>>(c0, B, (double)N * VI_EPS, covE, covD));
  else
    VI_KERNEL(VI_K_COV, st, k_eigvec<false><<<(unsigned)nc, (N + 31) / 32 * 32, smem_e, st>>>(c0, B, (double)N * VI_EPS, covE, covD));
  dim3 grid(tiles, tiles, (unsigned)nc);
  VI_KERNEL(VI_K_COV, st, k_bgemm<true><<<grid, 256, 0, st>>>(covE, NN, covE, NN, nullptr, covD, covH, NN, N));
  VI_KERNEL(VI_K_COV, st, k_bgemm<false><<<grid, 256, 0, st>>>(covH, NN, G, NN, B.rec + c0, nullptr, covT, NN, N));
  VI_KERNEL(VI_K_COV, st, k_bgemm<false><<<grid, 256, 0, st>>>(covT, NN, covH, NN, nullptr, nullptr, out, NN, N));
  VI_KERNEL(VI_K_COV, st, k_cov_nan<<<(unsigned)nc, 256, 0, st>>>(c0, N, B, out));
  return VI_OK;
}

}  // namespace

extern "C" int vi_fit_workspace_bytes(int32_t R, int32_t P, int32_t N, int32_t nreg, int64_t systems, int64_t* bytes) {
  VI_REQUIRE(bytes != nullptr && R >= 0 && N >= 1 && nreg >= 0, "bad arguments");
  if (N > 1024) { vi_set_error("nbasis %d > 1024 not supported", N); return VI_EUNSUPPORTED; }
  int64_t U = (int64_t)R * (nreg > 0 ? nreg : 1);
  int64_t wanted = U * VI_NALPHA;
  if (wanted < R) wanted = R;
  int64_t cap = systems > 0 ? vi_align_up(systems, 32) : default_system_cap(wanted, N, nreg, P);
  Bump b{nullptr, 0, 0};
  UnitBuf Ub;
  unit_carve(b, Ub, U);
  b.take<double>(VI_NALPHA);
  *bytes = b.off + per_system_bytes(N, nreg, P) * (cap + 32) + 16384 + cov_scratch_bytes(R, N) + gcv_scratch_bytes(R, P, U);
  return VI_OK;
}

// cap actually available inside a given workspace
static int64_t cap_for_workspace(int64_t ws_bytes, int64_t U, int n, int nreg, int P, int64_t Rcov) {
  Bump b{nullptr, 0, 0};
  UnitBuf Ub;
  unit_carve(b, Ub, U);
  b.take<double>(VI_NALPHA);
  int64_t left = ws_bytes - b.off - 8192 - (Rcov > 0 ? cov_scratch_bytes(Rcov, n) : 0);
  int64_t per = per_system_bytes(n, nreg, P);
  int64_t cap = left / per;
  cap = cap / 32 * 32;
  const int64_t wave = ql_wave(n);
  if (wave >= 32 && cap > wave) cap = cap / wave * wave;     // whole QL waves, as in default_system_cap
  return cap;
}

extern "C" int vi_solve_cov_batched(const double* G, const double* y, const int32_t* rec, const double* regmats,
                                    const double* lam, int64_t S, int32_t N, int32_t nreg, double rcond,
                                    double* C, double* dC, int32_t* rank, int32_t* status,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  VI_REQUIRE(G && y && C && rank && status && workspace, "NULL argument");
  VI_REQUIRE(S >= 0 && N >= 1 && N <= 1024 && nreg >= 0, "bad shape");
  VI_REQUIRE(nreg == 0 || (regmats && lam), "regmats/lam missing");
  if (S == 0) return VI_OK;
  cudaStream_t st = vi_stream(stream);
  const int64_t cc = S < kCovChunk ? S : kCovChunk;
  int64_t cap = cap_for_workspace(workspace_bytes, 0, N, nreg, 1, dC ? S : 0);
  if (cap < 32) { vi_set_error("workspace too small (%lld bytes)", (long long)workspace_bytes); return VI_EWORKSPACE; }
  if (cap > vi_align_up(S, 32)) cap = vi_align_up(S, 32);
  Bump b{reinterpret_cast<char*>(workspace), 0, workspace_bytes};
  double *covE = nullptr, *covH = nullptr, *covT = nullptr, *covD = nullptr;
  if (dC) {
    covE = b.take<double>(cc * (int64_t)N * N);
    covH = b.take<double>(cc * (int64_t)N * N);
    covT = b.take<double>(cc * (int64_t)N * N);
    covD = b.take<double>(cc * (int64_t)N);
  }
  SysBuf B;
  sysbuf_carve(b, B, cap, N, nreg, 1);
  VI_CUDA(cudaMemsetAsync(B.rot_total, 0, 4 * sizeof(unsigned long long), st));
  for (int64_t s0 = 0; s0 < S; s0 += cap) {
    int64_t cnt = (S - s0 < cap) ? S - s0 : cap;
    VI_KERNEL(VI_K_MISC, st, k_setup_solve<<<blocks(cap, 256), 256, 0, st>>>(s0, cnt, nreg, rec, lam, B));
    VI_LAUNCH_CHECK();
    if (int rc = run_systems(cnt, G, y, regmats, B, rcond, C + s0 * N, B.rank, st)) return rc;
    if (dC)
      for (int64_t c0 = 0; c0 < cnt; c0 += cc) {
        const int64_t nc = (cnt - c0 < cc) ? cnt - c0 : cc;
        if (int rc = run_cov_chunk(c0, nc, N, B, G, covE, covH, covT, covD, dC + (s0 + c0) * (int64_t)N * N, st)) return rc;
      }
    VI_KERNEL(VI_K_MISC, st, k_copy_status<<<blocks(cnt, 256), 256, 0, st>>>(cnt, B, rank + s0, status + s0));
    VI_LAUNCH_CHECK();
  }
  return VI_OK;
}

extern "C" int vi_solve_batched(const double* G, const double* y, const int32_t* rec, const double* regmats,
                                const double* lam, int64_t S, int32_t N, int32_t nreg, double rcond,
                                double* C, int32_t* rank, int32_t* status,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  return vi_solve_cov_batched(G, y, rec, regmats, lam, S, N, nreg, rcond, C, nullptr, rank, status, workspace,
                              workspace_bytes, stream);
}

extern "C" int vi_fit_batched(const double* At, const double* A, const double* Wm, const double* bm,
                              const double* G, const double* y, const int32_t* npts,
                              int32_t R, int32_t P, int32_t N,
                              const double* regmats, int32_t nreg, int32_t method,
                              double* C, double* dC, double* chi2, double* lam, int32_t* rank, int32_t* status,
                              int64_t* nsolve, void* workspace, int64_t workspace_bytes, void* stream) {
  VI_REQUIRE(At && Wm && bm && G && y && npts && C && chi2 && rank && status && workspace, "NULL argument");
  VI_REQUIRE(R >= 0 && P >= 1 && N >= 1 && N <= 1024 && nreg >= 0, "bad shape");
  VI_REQUIRE(method == VI_METHOD_NONE || method == VI_METHOD_CHI2 || method == VI_METHOD_GCV, "unknown method %d", method);
  VI_REQUIRE(method == VI_METHOD_NONE || (nreg >= 1 && regmats && lam), "chi2 / gcv need regularisation matrices");
  VI_REQUIRE(method != VI_METHOD_GCV || A != nullptr, "gcv needs the row-major design matrix A");
  if (nsolve) *nsolve = 0;
  if (R == 0) return VI_OK;
  if (method == VI_METHOD_NONE) nreg = 0;
  cudaStream_t st = vi_stream(stream);
  const double rcond = VI_EPS;
  const int64_t U = (int64_t)R * (nreg > 0 ? nreg : 1);

  const int64_t gcv_bytes = gcv_scratch_bytes(R, P, U);
  int64_t cap = cap_for_workspace(workspace_bytes - gcv_bytes, U, N, nreg, P, R);
  if (cap < 32) { vi_set_error("workspace too small (%lld bytes)", (long long)workspace_bytes); return VI_EWORKSPACE; }
  int64_t most = vi_align_up(method == VI_METHOD_CHI2 ? U * VI_NALPHA : (method == VI_METHOD_GCV ? U * (int64_t)P : (int64_t)R), 32);
  Bump b{reinterpret_cast<char*>(workspace), 0, workspace_bytes};
  int32_t* validx = b.take<int32_t>((int64_t)R * P);
  int32_t* nvalid = b.take<int32_t>(R);
  int64_t* ucount = b.take<int64_t>(U);
  UnitBuf Ub;
  unit_carve(b, Ub, U);
  double* pow10tab = b.take<double>(VI_NALPHA);
  const int64_t cc = R < kCovChunk ? R : kCovChunk;
  double* covE = b.take<double>(cc * (int64_t)N * N);
  double* covH = b.take<double>(cc * (int64_t)N * N);
  double* covT = b.take<double>(cc * (int64_t)N * N);
  double* covD = b.take<double>(cc * (int64_t)N);
  double* covRing = b.take<double>(2 * cc * (int64_t)N * N);
  if (cap > most) cap = most;
  SysBuf B;
  sysbuf_carve(b, B, cap, N, nreg, P);
  int64_t solved = 0;
  VI_CUDA(cudaMemsetAsync(B.rot_total, 0, 4 * sizeof(unsigned long long), st));

  if (method == VI_METHOD_CHI2) {
    double h_tab[VI_NALPHA];
    for (int k = 0; k < VI_NALPHA; ++k) h_tab[k] = pow(10.0, -(double)k);   // np.power(10., alpha), interpolate.py:250
    VI_CUDA(cudaMemcpyAsync(pow10tab, h_tab, sizeof(h_tab), cudaMemcpyHostToDevice, st));
    VI_CUDA(cudaMemsetAsync(Ub.tabbad, 0, U * sizeof(int32_t), st));
    VI_CUDA(cudaMemsetAsync(Ub.table, 0xff, U * VI_NALPHA * sizeof(double), st));   // unevaluated entries read as NaN (vi_fit_search_trace)
    VI_CUDA(cudaMemsetAsync(Ub.count, 0, 8 * sizeof(int32_t), st));
    // ---- phase 1: chi2(10^-k) table for every unit -------------------------------------
    VI_KERNEL(VI_K_MISC, st, k_kstar<<<(unsigned)U, 256, 0, st>>>(nreg, N, G, regmats, pow10tab, Ub.kstar));
    VI_KERNEL(VI_K_MISC, st, k_table_init<<<blocks(U, 128), 128, 0, st>>>(U, nreg, npts, Ub));
    // decades per pass: small when there are many units (little overshoot past the sign change, the
    // batches stay large), larger for few units (fewer latency-bound passes)
    int step = (int)(cap / (U > 0 ? U : 1));
    if (step < 4) step = 4;
    if (step > 16) step = 16;
    if (const char* e = getenv("VI_TABLE_STEP")) { int v = atoi(e); if (v >= 1) step = v; }
    // (A two-stream variant that overlapped apply/chi2 of one chunk with the tridiagonalisation of the next
    // was measured on B200 and gave no gain: 3.72 s vs 3.69 s per 10 k records; kept simple instead.)
    for (int pass = 0; pass < VI_NALPHA + 1; ++pass) {
      VI_KERNEL(VI_K_MISC, st, k_table_plan<<<blocks(U, 128), 128, 0, st>>>(U, step, Ub, ucount));
      VI_KERNEL(VI_K_MISC, st, k_scan_counts<<<1, 1024, 0, st>>>(U, ucount, Ub.off));
      int64_t T = 0;
      VI_CUDA(cudaMemcpyAsync(&T, Ub.off + U, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      VI_CUDA(cudaStreamSynchronize(st));
      if (T == 0) break;
      for (int64_t t0 = 0; t0 < T; t0 += cap) {
        int64_t cnt = (T - t0 < cap) ? T - t0 : cap;
        VI_KERNEL(VI_K_MISC, st, k_setup_table<<<blocks(cap, 256), 256, 0, st>>>(t0, cnt, U, nreg, pow10tab, Ub.off, Ub.kdone, B));
        if (int rc = run_systems(cnt, G, y, regmats, B, rcond, B.Csys, B.rank, st)) return rc;
        if (int rc = run_chi2(cnt, At, Wm, bm, P, B, B.Csys, B.chi2, st)) return rc;
        VI_KERNEL(VI_K_MISC, st, k_scatter_table<<<blocks(cnt, 256), 256, 0, st>>>(cnt, B, Ub));
      }
      solved += T;
      VI_KERNEL(VI_K_MISC, st, k_table_advance<<<blocks(U, 128), 128, 0, st>>>(U, nreg, npts, ucount, Ub));
    }
    // ---- phase 2: bracket + Brent in lock step ------------------------------------------
    VI_KERNEL(VI_K_MISC, st, k_bracket<<<blocks(U, 128), 128, 0, st>>>(U, nreg, npts, Ub));
    VI_LAUNCH_CHECK();
    const bool dbg = getenv("VI_DEBUG_ROUNDS") != nullptr;
    struct timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    for (int it = 0; it < VI_BRENT_MAXITER + 2; ++it) {
      int64_t round_total = 0;
      for (int64_t u0 = 0; u0 < U; u0 += cap) {
        int64_t cnt = (U - u0 < cap) ? U - u0 : cap;
        VI_CUDA(cudaMemsetAsync(Ub.count, 0, sizeof(int32_t), st));
        VI_KERNEL(VI_K_MISC, st, k_brent_propose<<<blocks(cnt, 128), 128, 0, st>>>(u0, cnt, nreg, B, Ub));
        int h_count = 0;
        VI_CUDA(cudaMemcpyAsync(&h_count, Ub.count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VI_CUDA(cudaStreamSynchronize(st));
        if (h_count == 0) continue;
        round_total += h_count;
        if (int rc = run_systems(h_count, G, y, regmats, B, rcond, B.Csys, B.rank, st)) return rc;
        if (int rc = run_chi2(h_count, At, Wm, bm, P, B, B.Csys, B.chi2, st)) return rc;
        VI_KERNEL(VI_K_MISC, st, k_brent_feed<<<blocks(h_count, 128), 128, 0, st>>>(h_count, B, Ub));
      }
      if (dbg) {
        struct timespec ts1;
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        fprintf(stderr, "[vi] brent round %d: %lld systems, %.2f ms since the previous round's count\n", it,
                (long long)round_total, (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
        ts0 = ts1;
      }
      if (round_total == 0) break;
      solved += round_total;
    }
  }
  if (method == VI_METHOD_GCV) {
    // ---- GCV: Nelder-Mead on alpha in lock step; one objective evaluation = one leave-one-gate-out
    // system per weighted gate of the record (interpolate.py:332-351) -----------------------------
    const Downdate dd{A, Wm, bm, P};
    VI_KERNEL(VI_K_MISC, st, k_valid_index<<<(unsigned)R, 32, 0, st>>>(P, Wm, validx, nvalid));
    VI_KERNEL(VI_K_MISC, st, k_nm_init<<<blocks(U, 128), 128, 0, st>>>(U, nreg, nvalid, Ub));
    for (int it = 0; it < 4 * VI_NM_MAXFUN + 8; ++it) {
      VI_KERNEL(VI_K_MISC, st, k_nm_propose<<<blocks(U, 128), 128, 0, st>>>(U, nreg, nvalid, Ub, ucount));
      VI_KERNEL(VI_K_MISC, st, k_scan_counts<<<1, 1024, 0, st>>>(U, ucount, Ub.off));
      int64_t T = 0;
      VI_CUDA(cudaMemcpyAsync(&T, Ub.off + U, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      VI_CUDA(cudaStreamSynchronize(st));
      if (T == 0) break;       // every unit has terminated (a unit with zero weighted gates never starts)
      for (int64_t t0 = 0; t0 < T; t0 += cap) {
        int64_t cnt = (T - t0 < cap) ? T - t0 : cap;
        VI_KERNEL(VI_K_MISC, st, k_gcv_setup<<<blocks(cap, 256), 256, 0, st>>>(t0, cnt, U, nreg, P, Ub.off, validx, B, Ub));
        if (int rc = run_tridiag(cnt, G, y, regmats, B, st, dd)) return rc;
        if (int rc = run_post(cnt, B, rcond, B.Csys, B.rank, st)) return rc;
        VI_KERNEL(VI_K_CHI2, st, k_gcv_resid<<<blocks(cnt, 8), 256, 0, st>>>(cnt, N, dd, B, B.Csys, B.chi2));
        VI_KERNEL(VI_K_MISC, st, k_gcv_accumulate<<<blocks(U, 128), 128, 0, st>>>(U, t0, cnt, Ub.off, B.chi2, Ub));
      }
      solved += T;
      VI_KERNEL(VI_K_MISC, st, k_nm_feed<<<blocks(U, 128), 128, 0, st>>>(U, Ub));
    }
  }
  // ---- phase 3: final solve with the found parameters (interpolate.py:566-569) -----------
  // dC may be a device pointer or a PINNED HOST pointer: in the second case every chunk of kCovChunk records is
  // computed into a two-slot device ring and copied out on a side stream while the next chunk is computed, so the
  // R x N x N block (1.66 GB at 10 k records, N = 144) neither occupies HBM nor serialises behind the fit.
  bool cov_to_host = false;
  int64_t cov_chunk = 0;
  CovSink& sink = cov_sink();
  if (dC != nullptr) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, dC) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
      cov_to_host = true;
      if (int rc = sink.init()) return rc;
    } else {
      (void)cudaGetLastError();
    }
  }
  for (int64_t r0 = 0; r0 < R; r0 += cap) {
    int64_t cnt = (R - r0 < cap) ? R - r0 : cap;
    VI_KERNEL(VI_K_MISC, st, k_setup_final<<<blocks(cap, 128), 128, 0, st>>>(r0, cnt, nreg, method, npts, B, Ub, lam, status));
    VI_LAUNCH_CHECK();
    if (int rc = run_systems(cnt, G, y, regmats, B, rcond, C + r0 * N, B.rank, st)) return rc;
    if (int rc = run_chi2(cnt, At, Wm, bm, P, B, C + r0 * N, B.chi2, st)) return rc;
    if (dC != nullptr) {
      const int64_t NN = (int64_t)N * N;
      for (int64_t c0 = 0; c0 < cnt; c0 += cc) {
        const int64_t nc = (cnt - c0 < cc) ? cnt - c0 : cc;
        double* dst = dC + (r0 + c0) * NN;
        double* out = dst;
        const int slot = (int)(cov_chunk & 1);
        if (cov_to_host) {        // chunk goes to the ring; wait until the copy that last used this slot has drained
          out = covRing + (int64_t)slot * cc * NN;
          if (cov_chunk >= 2) VI_CUDA(cudaStreamWaitEvent(st, sink.drained[slot], 0));
        }
        if (int rc = run_cov_chunk(c0, nc, N, B, G, covE, covH, covT, covD, out, st)) return rc;
        if (cov_to_host) {
          VI_CUDA(cudaEventRecord(sink.ready[slot], st));
          VI_CUDA(cudaStreamWaitEvent(sink.copy, sink.ready[slot], 0));
          VI_CUDA(cudaMemcpyAsync(dst, out, nc * NN * sizeof(double), cudaMemcpyDeviceToHost, sink.copy));
          VI_CUDA(cudaEventRecord(sink.drained[slot], sink.copy));
        }
        ++cov_chunk;
      }
    }
    VI_KERNEL(VI_K_MISC, st, k_finalize<<<(unsigned)cnt, 64, 0, st>>>(r0, cnt, N, B, C, chi2, rank, status));
    VI_LAUNCH_CHECK();
    solved += cnt;
  }
  {
    unsigned long long rot = 0;
    VI_CUDA(cudaMemcpyAsync(&rot, B.rot_total, sizeof(rot), cudaMemcpyDeviceToHost, st));
    VI_CUDA(cudaStreamSynchronize(st));
    vi_prof_count_rotations((int64_t)rot, solved);
  }
  if (cov_to_host) {        // the caller's stream order covers the copies: `st` waits for both slots
    for (int slot = 0; slot < 2 && slot < cov_chunk; ++slot) VI_CUDA(cudaStreamWaitEvent(st, sink.drained[slot], 0));
  }
  if (nsolve) *nsolve = solved;
  return VI_OK;
}

// Diagnostics of the last VI_METHOD_CHI2 search that ran in `workspace` (same R, P, nreg): the chi2(10^-k) table
// (entries the lazy walk never evaluated are NaN), nu = npts * scale factor of the bracket, the bracket decade and
// the number of distinct table entries evaluated.  This is what the parity report compares with the reference's
// own (alpha, chi2 - nu) trace (interpolate.py:180-207).
extern "C" int vi_fit_search_trace(const void* workspace, int64_t workspace_bytes, int32_t R, int32_t P, int32_t nreg,
                                   double* table, double* nu, int32_t* k_lo, int32_t* kdone, void* stream) {
  VI_REQUIRE(workspace && R >= 0 && P >= 1 && nreg >= 1, "bad arguments");
  if (R == 0) return VI_OK;
  cudaStream_t st = vi_stream(stream);
  const int64_t U = (int64_t)R * nreg;
  Bump b{reinterpret_cast<char*>(const_cast<void*>(workspace)), 0, workspace_bytes};
  b.take<int32_t>((int64_t)R * P);
  b.take<int32_t>(R);
  b.take<int64_t>(U);
  UnitBuf Ub;
  unit_carve(b, Ub, U);
  VI_REQUIRE(b.off <= workspace_bytes, "workspace smaller than the one the fit ran in");
  if (table) VI_CUDA(cudaMemcpyAsync(table, Ub.table, U * VI_NALPHA * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (nu) VI_CUDA(cudaMemcpyAsync(nu, Ub.nu, U * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (k_lo) VI_CUDA(cudaMemcpyAsync(k_lo, Ub.klo, U * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  if (kdone) VI_CUDA(cudaMemcpyAsync(kdone, Ub.kdone, U * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return VI_OK;
}

extern "C" int vi_fit_host(const double* A, const double* value, const double* error, const double* weight,
                           int32_t R, int32_t P, int32_t N, const double* regmats, int32_t nreg, int32_t method,
                           int32_t ne_mode, double* C, double* dC, double* chi2, double* lam, int32_t* rank,
                           int32_t* status) {
  VI_REQUIRE(A && value && error && C && chi2 && rank && status, "NULL argument");
  VI_REQUIRE(R >= 0 && P >= 1 && N >= 1, "bad shape");
  if (R == 0) return VI_OK;
  cudaStream_t s = nullptr;
  int64_t ws_bytes = 0;
  if (int rc = vi_fit_workspace_bytes(R, P, N, nreg, 0, &ws_bytes)) return rc;
  const size_t RP = (size_t)R * P, PN = (size_t)P * N, NN = (size_t)N * N;
  std::vector<void*> owned;
  auto dalloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    owned.push_back(p);
    return p;
  };
  auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
  double* dA = (double*)dalloc(PN * 8);
  double* dAt = (double*)dalloc(PN * 8);
  double* dval = (double*)dalloc(RP * 8);
  double* derr = (double*)dalloc(RP * 8);
  double* dwt = weight ? (double*)dalloc(RP * 8) : nullptr;
  double* dWm = (double*)dalloc(RP * 8);
  double* dbm = (double*)dalloc(RP * 8);
  double* dG = (double*)dalloc((size_t)R * NN * 8);
  double* dy = (double*)dalloc((size_t)R * N * 8);
  double* dreg = (double*)dalloc((size_t)(nreg > 0 ? nreg : 1) * NN * 8);
  double* dC_ = (double*)dalloc((size_t)R * N * 8);
  double* ddC = dC ? (double*)dalloc((size_t)R * NN * 8) : nullptr;
  double* dchi = (double*)dalloc((size_t)R * 8);
  double* dlam = (double*)dalloc((size_t)R * (nreg > 0 ? nreg : 1) * 8);
  int32_t* dnp = (int32_t*)dalloc((size_t)R * 4);
  int32_t* drank = (int32_t*)dalloc((size_t)R * 4);
  int32_t* dst = (int32_t*)dalloc((size_t)R * 4);
  void* ws = dalloc((size_t)ws_bytes);
  if (!dA || !dAt || !dval || !derr || !dWm || !dbm || !dG || !dy || !dreg || !dC_ || !dchi || !dlam || !dnp ||
      !drank || !dst || !ws || (weight && !dwt) || (dC && !ddC)) {
    cleanup();
    vi_set_error("cudaMalloc failed (workspace %lld bytes)", (long long)ws_bytes);
    return VI_ECUDA;
  }
  int rc = VI_OK;
#define VI_TRY(x) do { if (rc == VI_OK) { cudaError_t _e = (x); if (_e != cudaSuccess) { vi_set_error("%s -> %s", #x, cudaGetErrorString(_e)); rc = VI_ECUDA; } } } while (0)
  VI_TRY(cudaMemcpyAsync(dA, A, PN * 8, cudaMemcpyHostToDevice, s));
  VI_TRY(cudaMemcpyAsync(dval, value, RP * 8, cudaMemcpyHostToDevice, s));
  VI_TRY(cudaMemcpyAsync(derr, error, RP * 8, cudaMemcpyHostToDevice, s));
  if (weight) VI_TRY(cudaMemcpyAsync(dwt, weight, RP * 8, cudaMemcpyHostToDevice, s));
  if (nreg > 0) VI_TRY(cudaMemcpyAsync(dreg, regmats, (size_t)nreg * NN * 8, cudaMemcpyHostToDevice, s));
  if (rc == VI_OK) {
    // (not VI_KERNEL: its launch check returns, which would skip cleanup())
    vi_prof_launch_begin(VI_K_MISC, s);
    k_transpose<<<blocks((int64_t)PN, 256), 256, 0, s>>>(dA, P, N, dAt);
    vi_prof_launch_end(VI_K_MISC, s);
    VI_TRY(cudaGetLastError());
  }
  if (rc == VI_OK) {
    rc = vi_normal_eq_batched(dA, dval, derr, dwt, R, P, N, ne_mode, dG, dy, nullptr, dnp, dWm, dbm, s);
  }
  if (rc == VI_OK)
    rc = vi_fit_batched(dAt, dA, dWm, dbm, dG, dy, dnp, R, P, N, dreg, nreg, method, dC_, ddC, dchi, dlam, drank, dst,
                        nullptr, ws, ws_bytes, s);
  VI_TRY(cudaMemcpyAsync(C, dC_, (size_t)R * N * 8, cudaMemcpyDeviceToHost, s));
  if (dC) VI_TRY(cudaMemcpyAsync(dC, ddC, (size_t)R * NN * 8, cudaMemcpyDeviceToHost, s));
  VI_TRY(cudaMemcpyAsync(chi2, dchi, (size_t)R * 8, cudaMemcpyDeviceToHost, s));
  if (lam && nreg > 0) VI_TRY(cudaMemcpyAsync(lam, dlam, (size_t)R * nreg * 8, cudaMemcpyDeviceToHost, s));
  VI_TRY(cudaMemcpyAsync(rank, drank, (size_t)R * 4, cudaMemcpyDeviceToHost, s));
  VI_TRY(cudaMemcpyAsync(status, dst, (size_t)R * 4, cudaMemcpyDeviceToHost, s));
  VI_TRY(cudaStreamSynchronize(s));
#undef VI_TRY
  cleanup();
  return rc;
}
