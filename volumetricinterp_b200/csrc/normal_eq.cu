// K2 — error-weighted normal equations for many time records at once.
//
// Replaces, for R records per launch, the reference's
//   mask = isfinite(ne0)                         interpolate.py:516-520
//   W = er0**-2 ; b = ne0                        interpolate.py:523-524
//   AWA = einsum('ji,j,jk->ik', A, W, A)         interpolate.py:456
//   y   = einsum('ji,j,j->i',  A, W, b)          interpolate.py:458
// The design matrix A (P x N) is evaluated once for all P gates; invalid gates get w = 0,
// b = 0, which is bit-identical to the reference's row deletion (adding +0.0 terms).
//
// Two kernels:
//   k_ne_strict  VI_NE_STRICT: acc = (A[j,i]*w[j])*A[j,k] + acc, sequential in j, two
//                roundings per term, never fused — reproduces np.einsum bit for bit.
//   k_ne_dmma3   VI_NE_FAST: per record a symmetric rank-P update G = (A.w)^T A on the FP64
//                tensor-core path (mma.sync m16n8k16 / m8n8k4 f64), lower triangle only, the
//                right-hand side y = A^T (w.b) on the FP64 ALUs from the fragments the diagonal
//                units hold.  A is staged through shared memory by cp.async double buffering.
#include "common.cuh"
#include <stdlib.h>

namespace {

// correctly rounded 1/(e*e) (double-double product, one Newton correction)
__device__ __forceinline__ double inv_square(double e) {
  double hi = e * e;
  double lo = fma(e, e, -hi);
  double q = 1.0 / hi;
  double r = fma(-q, hi, 1.0);
  r = fma(-q, lo, r);
  return fma(r, q, q);
}

// masked weight / datum of gate idx (flattened r*P + j)
__device__ __forceinline__ void load_wb(const double* __restrict__ value, const double* __restrict__ error,
                                        const double* __restrict__ weight, int64_t idx, double& w, double& b) {
  double v = value[idx];
  bool ok = isfinite(v);
  double ww = weight ? weight[idx] : inv_square(error[idx]);
  w = ok ? ww : 0.0;
  b = ok ? v : 0.0;
}

__global__ void __launch_bounds__(256)
k_prep(const double* __restrict__ value, const double* __restrict__ error, const double* __restrict__ weight,
       int P, double* __restrict__ sWbb, int32_t* __restrict__ npts, double* __restrict__ Wm, double* __restrict__ bm) {
  __shared__ double ssum[256];
  __shared__ int scnt[256];
  const int r = blockIdx.x;
  double s = 0.0;
  int cnt = 0;
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    int64_t idx = (int64_t)r * P + j;
    double w, b;
    load_wb(value, error, weight, idx, w, b);
    if (isfinite(value[idx])) ++cnt;
    s += w * b * b;
    if (Wm) Wm[idx] = w;
    if (bm) bm[idx] = b;
  }
  ssum[threadIdx.x] = s;
  scnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { if (sWbb) sWbb[r] = ssum[0]; if (npts) npts[r] = scnt[0]; }
}

// ---------------------------------------------------------------------------------------------
// strict: 128 x 128 output super-block per CTA, 8 x 8 per thread, gates staged 32 at a time
// ---------------------------------------------------------------------------------------------
constexpr int kSB = 128, kST = 8, kSJ = 32;

__global__ void __launch_bounds__(256)
k_ne_strict(const double* __restrict__ A, const double* __restrict__ value, const double* __restrict__ error,
            const double* __restrict__ weight, int P, int N, int nsb, double* __restrict__ G, double* __restrict__ y) {
  extern __shared__ double sm[];
  double* Ai = sm;                    // kSJ x kSB
  double* Ak = sm + kSJ * kSB;        // kSJ x kSB
  double* sw = sm + 2 * kSJ * kSB;    // kSJ
  double* sb = sw + kSJ;              // kSJ
  const int r = blockIdx.x / (nsb * nsb);
  const int sbi = (blockIdx.x / nsb) % nsb, sbk = blockIdx.x % nsb;
  const int ti = threadIdx.x / 16, tk = threadIdx.x % 16;
  const int i0 = sbi * kSB, k0 = sbk * kSB;
  double acc[kST][kST];
  double yacc[kST];
#pragma unroll
  for (int a = 0; a < kST; ++a) {
    yacc[a] = 0.0;
#pragma unroll
    for (int b = 0; b < kST; ++b) acc[a][b] = 0.0;
  }
  const bool do_y = (sbk == 0) && (tk == 0);
  for (int j0 = 0; j0 < P; j0 += kSJ) {
    __syncthreads();
    for (int e = threadIdx.x; e < kSJ * kSB; e += 256) {
      int jj = e / kSB, c = e % kSB;
      int j = j0 + jj;
      double vi = 0.0, vk = 0.0;
      if (j < P) {
        if (i0 + c < N) vi = A[(int64_t)j * N + i0 + c];
        if (k0 + c < N) vk = A[(int64_t)j * N + k0 + c];
      }
      Ai[e] = vi;
      Ak[e] = vk;
    }
    if (threadIdx.x < kSJ) {
      int j = j0 + threadIdx.x;
      double w = 0.0, b = 0.0;
      if (j < P) load_wb(value, error, weight, (int64_t)r * P + j, w, b);
      sw[threadIdx.x] = w;
      sb[threadIdx.x] = b;
    }
    __syncthreads();
    const int jn = min(kSJ, P - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const double w = sw[jj];
      double t[kST], ak[kST];
#pragma unroll
      for (int a = 0; a < kST; ++a) t[a] = __dmul_rn(Ai[jj * kSB + ti + 16 * a], w);
#pragma unroll
      for (int b = 0; b < kST; ++b) ak[b] = Ak[jj * kSB + tk + 16 * b];
#pragma unroll
      for (int a = 0; a < kST; ++a)
#pragma unroll
        for (int b = 0; b < kST; ++b) acc[a][b] = __dadd_rn(__dmul_rn(t[a], ak[b]), acc[a][b]);
      if (do_y) {
        const double bj = sb[jj];
#pragma unroll
        for (int a = 0; a < kST; ++a) yacc[a] = __dadd_rn(__dmul_rn(t[a], bj), yacc[a]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < kST; ++a) {
    int i = i0 + ti + 16 * a;
    if (i >= N) continue;
#pragma unroll
    for (int b = 0; b < kST; ++b) {
      int k = k0 + tk + 16 * b;
      if (k < N) G[((int64_t)r * N + i) * N + k] = acc[a][b];
    }
    if (do_y) y[(int64_t)r * N + i] = yacc[a];
  }
}

// ---------------------------------------------------------------------------------------------
// fast: FP64 tensor-core (DMMA) symmetric rank-P update, one record per CTA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// fast: FP64 tensor-core (DMMA) symmetric rank-P update, one record per CTA (k_ne_dmma3 below).
// Staging shared by the kernel: A transposed in shared memory (kELD doubles per basis function and
// stage, gate index permuted so that the four k-values a lane feeds to one instruction are adjacent:
// every fragment is two 128-bit shared loads), masked weights / data streamed by cp.async from the
// arrays k_prep wrote (no division in the pipeline), one CTA barrier per stage.
// ---------------------------------------------------------------------------------------------
constexpr int kEJ = 80;          // gates per stage (five k16 steps)
constexpr int kELD = 82;         // doubles per staged column: 82 = 2 (mod 16) -> conflict-free 128-bit fragment loads
constexpr int kEStages = 2;

__device__ __forceinline__ void dmma_16x8x16(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// position of gate jj (0..kEJ-1) inside a staged column: within each block of 16 gates, k -> 4 (k % 4) + k / 4
__device__ __forceinline__ int gate_slot(int jj) { return (jj & ~15) | ((jj & 3) << 2) | ((jj >> 2) & 3); }

// ---------------------------------------------------------------------------------------------
// the kernel: only the work that is part of the lower triangle.
//   * the right-hand side y = A^T W b no longer occupies a 16x8 tile per row block (one useful column
//     of eight): it is accumulated with plain DFMAs from the A fragments the diagonal unit holds anyway;
//   * the tile right of the diagonal block's first half, (mi, 2 mi + 1), only has its lower 8 rows
//     below the diagonal: it is issued as four m8n8k4 (half the tensor work of an m16n8k16);
//   * units are dealt to the 12 warps by weight (full tile 4, diagonal unit 3), contiguously in
//     (mi, ni) order so a warp still reloads its A fragment only when mi changes.
// Issued tensor work per record: 81 full + 9 half tiles = 10 944 pairs for 10 440 needed (95 %),
// against 99 x 128 = 12 672 (82 %) in version 2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

#include "vi_ne_split.h"

template <int DT>
__global__ void __launch_bounds__(kW3 * 32)
k_ne_dmma3(const double* __restrict__ A, const double* __restrict__ Wm, const double* __restrict__ bm,
           int P, int N, int mt, int cols, const __grid_constant__ NeSplit sp, double* __restrict__ G,
           double* __restrict__ y) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                                              // kEStages x cols x kELD
  double* sw = sm + (size_t)kEStages * cols * kELD;            // kEStages x kEJ (slot order)
  double* sb = sw + kEStages * kEJ;                            // kEStages x kEJ (slot order)
  const int r = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int e = tid; e < kEStages * cols * kELD + 2 * kEStages * kEJ; e += blockDim.x) sm[e] = 0.0;
  __syncthreads();
  const int t0 = sp.wbeg[warp];
  const int t1 = sp.wend[warp];
  int tinfo[DT];
  double acc[DT][4];
#pragma unroll
  for (int q = 0; q < DT; ++q) {
    const int tt = t0 + q;
    tinfo[q] = (tt < t1) ? sp.uinfo[tt] : -1;
    acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
  }
  const int nchunk = (P + kEJ - 1) / kEJ;
  const int rj = warp, cc = lane;                  // a warp copies one gate (one row of A) per trip
  const double* Wr = Wm + (int64_t)r * P;
  const double* br = bm + (int64_t)r * P;

  auto stage_load = [&](int chunk) {
    const int st = chunk % kEStages;
    double* dst = S + (size_t)st * cols * kELD;
    const int j0 = chunk * kEJ;
#pragma unroll
    for (int q = 0; q < kEJ / kW3; ++q) {
      const int jj = rj + kW3 * q;
      const int j = j0 + jj;
      const int sl = gate_slot(jj);
      if (j < P) {
        const double* src = A + (int64_t)j * N;
        for (int c = cc; c < N; c += 32) cp_async8(dst + c * kELD + sl, src + c);
      } else {
        for (int c = cc; c < N; c += 32) dst[c * kELD + sl] = 0.0;
      }
    }
    if (tid < kEJ) {
      const int j = j0 + tid, sl = gate_slot(tid);
      if (j < P) { cp_async8(sw + st * kEJ + sl, Wr + j); cp_async8(sb + st * kEJ + sl, br + j); }
      else { sw[st * kEJ + sl] = 0.0; sb[st * kEJ + sl] = 0.0; }
    }
    cp_async_commit();
  };

  for (int c = 0; c < kEStages - 1 && c < nchunk; ++c) stage_load(c);
  for (int ch = 0; ch < nchunk; ++ch) {
    if (kEStages >= 3 && ch + 1 < nchunk) cp_async_wait<kEStages - 2>(); else cp_async_wait<0>();
    __syncthreads();
    if (ch + kEStages - 1 < nchunk) stage_load(ch + kEStages - 1);
    const int st = ch % kEStages;
    const double* Sst = S + (size_t)st * cols * kELD;
    const double* Wst = sw + st * kEJ;
    const double* Bst = sb + st * kEJ;
#pragma unroll
    for (int ks = 0; ks < kEJ / 16; ++ks) {
      const int kb = 16 * ks + 4 * t;
      const double2 w01 = *reinterpret_cast<const double2*>(Wst + kb);
      const double2 w23 = *reinterpret_cast<const double2*>(Wst + kb + 2);
      int cur_mi = -1;
      double a[8];
#pragma unroll
      for (int q = 0; q < DT; ++q) {
        if (tinfo[q] >= 0) {
          const int mi = tinfo[q] & 255, ni = (tinfo[q] >> 8) & 255;
          if (mi != cur_mi) {
            cur_mi = mi;
            const double* p0 = Sst + (16 * mi + g) * kELD + kb;
            const double* p1 = p0 + 8 * kELD;
            const double2 x01 = *reinterpret_cast<const double2*>(p0), x23 = *reinterpret_cast<const double2*>(p0 + 2);
            const double2 z01 = *reinterpret_cast<const double2*>(p1), z23 = *reinterpret_cast<const double2*>(p1 + 2);
            a[0] = x01.x * w01.x; a[2] = x01.y * w01.y; a[4] = x23.x * w23.x; a[6] = x23.y * w23.y;
            a[1] = z01.x * w01.x; a[3] = z01.y * w01.y; a[5] = z23.x * w23.x; a[7] = z23.y * w23.y;
          }
          if (!(tinfo[q] & kUnitDiag)) {
            const double* pb = Sst + (8 * ni + g) * kELD + kb;
            const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
            const double b[4] = {b01.x, b01.y, b23.x, b23.y};
            dmma_16x8x16(acc[q], a, b);
          } else {
            if (tinfo[q] & kUnitHalf) {
              const double* pb = Sst + (8 * ni + g) * kELD + kb;
              const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
              dmma_8x8x4(acc[q][2], acc[q][3], a[1], b01.x);
              dmma_8x8x4(acc[q][2], acc[q][3], a[3], b01.y);
              dmma_8x8x4(acc[q][2], acc[q][3], a[5], b23.x);
              dmma_8x8x4(acc[q][2], acc[q][3], a[7], b23.y);
            }
            // rhs: y_i += sum_k (A_ki w_k) b_k over this lane's four gates (rows g and g + 8 of the block)
            const double2 v01 = *reinterpret_cast<const double2*>(Bst + kb);
            const double2 v23 = *reinterpret_cast<const double2*>(Bst + kb + 2);
            acc[q][0] = fma(a[0], v01.x, acc[q][0]); acc[q][0] = fma(a[2], v01.y, acc[q][0]);
            acc[q][0] = fma(a[4], v23.x, acc[q][0]); acc[q][0] = fma(a[6], v23.y, acc[q][0]);
            acc[q][1] = fma(a[1], v01.x, acc[q][1]); acc[q][1] = fma(a[3], v01.y, acc[q][1]);
            acc[q][1] = fma(a[5], v23.x, acc[q][1]); acc[q][1] = fma(a[7], v23.y, acc[q][1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < DT; ++q) {
    if (tinfo[q] < 0) continue;
    const int mi = tinfo[q] & 255, ni = (tinfo[q] >> 8) & 255;
    const bool diag = tinfo[q] & kUnitDiag;
    if (diag) {
      // rhs: the four lanes of a group hold the partial sums of different gates
      double y0 = acc[q][0], y1 = acc[q][1];
      y0 += __shfl_xor_sync(0xffffffffu, y0, 1); y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
      y0 += __shfl_xor_sync(0xffffffffu, y0, 2); y1 += __shfl_xor_sync(0xffffffffu, y1, 2);
      if (t == 0) {
        const int i = 16 * mi + g;
        if (i < N) y[(int64_t)r * N + i] = y0;
        if (i + 8 < N) y[(int64_t)r * N + i + 8] = y1;
      }
      if (!(tinfo[q] & kUnitHalf)) continue;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (diag && v < 2) continue;
      const int i = 16 * mi + g + ((v & 2) ? 8 : 0);
      const int k = 8 * ni + 2 * t + (v & 1);
      if (i >= N) continue;
      const double val = acc[q][v];
      if (k <= i) {
        G[((int64_t)r * N + i) * N + k] = val;
        if (k != i) G[((int64_t)r * N + k) * N + i] = val;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// fast, orders beyond the one-CTA form (N > 160: the high-order model of BASELINE configs[2], N = 500; radbasfun
// NUMGRIDPNT = 7, N = 343): the lower triangle of G in 128 x 128 blocks, one CTA per (block, record), 8 warps of
// 32 x 64 outputs (16 m16n8k16 tiles each), both operands staged k-slot permuted from the rows of A (32 gates per
// stage, two stages, 8-byte cp.async: the transposition happens in the copy), weights applied to the A fragments,
// the right-hand side from the fragments of the diagonal blocks.  Diagonal blocks compute their upper half too
// (4 of 10 blocks at N = 500: 20 % extra tensor work) -- this is the secondary configuration.
// ---------------------------------------------------------------------------------------------
constexpr int kBG = 32;          // gates per stage
constexpr int kBLD = 34;         // doubles per staged column: 34 = 2 (mod 16)
constexpr int kBB = 128;         // block edge
__global__ void __launch_bounds__(256)
k_ne_dmma_blk(const double* __restrict__ A, const double* __restrict__ Wm, const double* __restrict__ bm,
              int P, int N, int nb, double* __restrict__ G, double* __restrict__ y) {
  extern __shared__ __align__(16) double sm[];
  // per stage: Si (kBB x kBLD), Sj (kBB x kBLD), w (kBG), b (kBG)
  const int stage_doubles = 2 * kBB * kBLD + 2 * kBG;
  const int r = blockIdx.y;
  int bi = 0, bj = 0;
  {                                                            // blockIdx.x -> (bi >= bj)
    int rem = blockIdx.x;
    while (rem > bi) { rem -= bi + 1; ++bi; }
    bj = rem;
  }
  const bool diag = bi == bj;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wr = warp & 3, wc = warp >> 2;                     // 32-row group, 64-column half
  const double* Wr = Wm + (int64_t)r * P;
  const double* br = bm + (int64_t)r * P;
  const int nchunk = (P + kBG - 1) / kBG;
  auto stage_load = [&](int chunk) {
    double* Si = sm + (size_t)(chunk & 1) * stage_doubles;
    double* Sj = Si + kBB * kBLD;
    double* sw = Sj + kBB * kBLD;
    double* sb = sw + kBG;
    const int j0 = chunk * kBG;
#pragma unroll
    for (int q = 0; q < kBG / 8; ++q) {                        // a warp copies one gate (one row of A) per trip
      const int jj = warp + 8 * q, j = j0 + jj;
      const int sl = gate_slot(jj);
      const double* src = A + (int64_t)j * N;
#pragma unroll
      for (int c4 = 0; c4 < kBB / 32; ++c4) {
        const int c = lane + 32 * c4;
        const int ci = kBB * bi + c, cj = kBB * bj + c;
        if (j < P && ci < N) cp_async8(Si + c * kBLD + sl, src + ci); else Si[c * kBLD + sl] = 0.0;
        if (!diag) { if (j < P && cj < N) cp_async8(Sj + c * kBLD + sl, src + cj); else Sj[c * kBLD + sl] = 0.0; }
      }
    }
    if (tid < kBG) {
      const int j = j0 + tid, sl = gate_slot(tid);
      if (j < P) { cp_async8(sw + sl, Wr + j); cp_async8(sb + sl, br + j); }
      else { sw[sl] = 0.0; sb[sl] = 0.0; }
    }
    cp_async_commit();
  };
  double acc[2][8][4];
  double ya[2][2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    ya[m][0] = ya[m][1] = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[m][q][0] = acc[m][q][1] = acc[m][q][2] = acc[m][q][3] = 0.0;
  }
  stage_load(0);
  for (int ch = 0; ch < nchunk; ++ch) {
    if (ch + 1 < nchunk) { stage_load(ch + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const double* Si = sm + (size_t)(ch & 1) * stage_doubles;
    const double* Sj = diag ? Si : Si + kBB * kBLD;
    const double* sw = Si + 2 * kBB * kBLD;
    const double* sb = sw + kBG;
#pragma unroll
    for (int ks = 0; ks < kBG / 16; ++ks) {
      const int kb = 16 * ks + 4 * t;
      const double2 w01 = *reinterpret_cast<const double2*>(sw + kb);
      const double2 w23 = *reinterpret_cast<const double2*>(sw + kb + 2);
      double a[2][8];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const double* p0 = Si + (32 * wr + 16 * m + g) * kBLD + kb;
        const double* p1 = p0 + 8 * kBLD;
        const double2 x01 = *reinterpret_cast<const double2*>(p0), x23 = *reinterpret_cast<const double2*>(p0 + 2);
        const double2 z01 = *reinterpret_cast<const double2*>(p1), z23 = *reinterpret_cast<const double2*>(p1 + 2);
        a[m][0] = x01.x * w01.x; a[m][2] = x01.y * w01.y; a[m][4] = x23.x * w23.x; a[m][6] = x23.y * w23.y;
        a[m][1] = z01.x * w01.x; a[m][3] = z01.y * w01.y; a[m][5] = z23.x * w23.x; a[m][7] = z23.y * w23.y;
      }
      if (diag && wc == 0) {                                   // rhs: y_i += sum_k (A_ki w_k) b_k
        const double2 v01 = *reinterpret_cast<const double2*>(sb + kb);
        const double2 v23 = *reinterpret_cast<const double2*>(sb + kb + 2);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          ya[m][0] = fma(a[m][0], v01.x, ya[m][0]); ya[m][0] = fma(a[m][2], v01.y, ya[m][0]);
          ya[m][0] = fma(a[m][4], v23.x, ya[m][0]); ya[m][0] = fma(a[m][6], v23.y, ya[m][0]);
          ya[m][1] = fma(a[m][1], v01.x, ya[m][1]); ya[m][1] = fma(a[m][3], v01.y, ya[m][1]);
          ya[m][1] = fma(a[m][5], v23.x, ya[m][1]); ya[m][1] = fma(a[m][7], v23.y, ya[m][1]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double* pb = Sj + (64 * wc + 8 * q + g) * kBLD + kb;
        const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
        dmma_16x8x16(acc[0][q], a[0], b);
        dmma_16x8x16(acc[1][q], a[1], b);
      }
    }
    __syncthreads();
  }
  // c0 (row g, col 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
  double* Gr = G + (int64_t)r * N * N;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = kBB * bi + 32 * wr + 16 * m + g + ((v & 2) ? 8 : 0);
        const int k = kBB * bj + 64 * wc + 8 * q + 2 * t + (v & 1);
        if (i < N && k <= i) {
          Gr[(int64_t)i * N + k] = acc[m][q][v];
          if (k != i) Gr[(int64_t)k * N + i] = acc[m][q][v];
        }
      }
  if (diag && wc == 0) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      double y0 = ya[m][0], y1 = ya[m][1];
      y0 += __shfl_xor_sync(0xffffffffu, y0, 1); y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
      y0 += __shfl_xor_sync(0xffffffffu, y0, 2); y1 += __shfl_xor_sync(0xffffffffu, y1, 2);
      if (t == 0) {
        const int i = kBB * bi + 32 * wr + 16 * m + g;
        if (i < N) y[(int64_t)r * N + i] = y0;
        if (i + 8 < N) y[(int64_t)r * N + i + 8] = y1;
      }
    }
  }
}

}  // namespace

extern "C" int vi_normal_eq_batched(const double* A, const double* value, const double* error, const double* weight,
                                    int32_t R, int32_t P, int32_t N, int32_t mode, double* G, double* y,
                                    double* sWbb, int32_t* npts, double* Wm, double* bm, void* stream) {
  VI_REQUIRE(A && value && (error || weight) && G && y, "NULL argument");
  VI_REQUIRE(R >= 0 && P >= 1 && N >= 1, "bad shape R=%d P=%d N=%d", R, P, N);
  if (R == 0) return VI_OK;
  cudaStream_t s = vi_stream(stream);
  if (sWbb || npts || Wm || bm) {
    VI_KERNEL(VI_K_NORMAL_EQ, s, k_prep<<<R, 256, 0, s>>>(value, error, weight, P, sWbb, npts, Wm, bm));
    VI_LAUNCH_CHECK();
  }
  auto strict = [&]() -> int {
    int nsb = (N + kSB - 1) / kSB;
    size_t smem = (size_t)(2 * kSJ * kSB + 2 * kSJ) * sizeof(double);
    VI_CUDA(cudaFuncSetAttribute(k_ne_strict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VI_KERNEL(VI_K_NORMAL_EQ, s, k_ne_strict<<<(unsigned)(R * nsb * nsb), 256, smem, s>>>(A, value, error, weight, P, N, nsb, G, y));
    VI_LAUNCH_CHECK();
    return VI_OK;
  };
  if (mode == VI_NE_STRICT) return strict();
  if (mode != VI_NE_FAST) { vi_set_error("unknown normal-equation mode %d", mode); return VI_EINVAL; }
  // the tensor-core kernel keeps one record's whole lower triangle of 16 x 8 tiles in one CTA (N <= 160); larger
  // models (the reference has no limit: radbasfun NUMGRIDPNT = 7 is N = 343, interpolate.py:456) run the tiled
  // strict kernel
  const int mt = (N + 15) / 16;
  const int cols3 = 16 * mt;
  const size_t smem3 = ((size_t)kEStages * cols3 * kELD + 2 * kEStages * kEJ) * sizeof(double) + 64;
  NeSplit sp;
  const bool one_cta = smem3 <= 227 * 1024 && ne3_split(N, mt, 7, sp);
  static const bool no_blk = getenv("VI_NE_BIG_STRICT") != nullptr;      // A/B: strict kernel for the large orders
  if (!one_cta && (no_blk || R > 65535)) return strict();
  // masked weights / data: the caller's arrays, else a stream-ordered temporary
  double *Wt = nullptr, *bt = nullptr;
  if (!Wm || !bm) {
    VI_CUDA(cudaMallocAsync(&Wt, (size_t)R * P * sizeof(double), s));
    if (cudaMallocAsync(&bt, (size_t)R * P * sizeof(double), s) != cudaSuccess) {
      cudaFreeAsync(Wt, s);
      vi_set_error("out of device memory for the masked weights (%lld bytes)", (long long)R * P * 8);
      return VI_ECUDA;
    }
    VI_KERNEL(VI_K_NORMAL_EQ, s, k_prep<<<R, 256, 0, s>>>(value, error, weight, P, nullptr, nullptr, Wt, bt));
    Wm = Wt; bm = bt;
  }
  cudaError_t e;
  if (one_cta) {
    e = cudaFuncSetAttribute(k_ne_dmma3<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
    if (e == cudaSuccess) {
      vi_prof_launch_begin(VI_K_NORMAL_EQ, s);
      k_ne_dmma3<7><<<(unsigned)R, kW3 * 32, smem3, s>>>(A, Wm, bm, P, N, mt, cols3, sp, G, y);
      vi_prof_launch_end(VI_K_NORMAL_EQ, s);
      e = cudaGetLastError();
    }
  } else {
    const int nb = (N + kBB - 1) / kBB;
    const size_t smemb = (size_t)2 * (2 * kBB * kBLD + 2 * kBG) * sizeof(double);
    e = cudaFuncSetAttribute(k_ne_dmma_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemb);
    if (e == cudaSuccess) {
      vi_prof_launch_begin(VI_K_NORMAL_EQ, s);
      k_ne_dmma_blk<<<dim3((unsigned)(nb * (nb + 1) / 2), (unsigned)R), 256, smemb, s>>>(A, Wm, bm, P, N, nb, G, y);
      vi_prof_launch_end(VI_K_NORMAL_EQ, s);
      e = cudaGetLastError();
    }
  }
  if (Wt) { cudaFreeAsync(Wt, s); cudaFreeAsync(bt, s); }
  if (e != cudaSuccess) { vi_set_error("tensor-core normal equations: %s", cudaGetErrorString(e)); return VI_ECUDA; }
  return VI_OK;
}
