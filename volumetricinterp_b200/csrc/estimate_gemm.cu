// K4, many records: out[r][p] = sum_n basis(p)[n] C[r][n]  (estimate.py:113-115) as an FP64 tensor-core GEMM
// [points x N] . [N x records], for the points INSIDE the hull only (estimate.py:119-121 turns the others into NaN,
// so neither their basis rows nor their GEMM rows are ever needed).
//
// Pipeline of vi_estimate_*_many (host side at the bottom):
//   k_hull_compact   (basis.cu)  in-hull test per point + compaction: idx[0 .. count) = the points to evaluate
//   k_rows_*_idx     (basis.cu)  one thread per compacted point: its basis row -> Arows[slot(n)][j] (full occupancy:
//                                the special-function work no longer sits in front of the MMAs of a 4-warp CTA)
//   k_coef_slots     (here)      C -> slot order, zero padded to a multiple of 16 columns
//   k_fill_nan       (here)      the whole output tile = NaN (the GEMM overwrites the in-hull points); on a side
//                                stream, beside the two kernels above
//   k_est_gemm       (here)      persistent CTAs over (128 compacted points x 32-record chunk) units, 8 warps, mma.sync.m16n8k16.f64
//                                (SASS DMMA.8x8x4), operands staged by 16-byte cp.async, double buffered
// Both operands are stored with the k index permuted inside each group of 16 (k -> 4 (k % 4) + k / 4) and a row
// stride = 2 (mod 16) doubles, so that every fragment is two conflict-free 128-bit shared loads.
#include "common.cuh"
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int kGemmThreads = 256;
constexpr int kTileP = 128;        // compacted points per CTA
constexpr int kRC = 32;            // records per chunk

__host__ __device__ inline int k_slot(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }

__device__ __forceinline__ void dmma_16x8x16(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_)); }

__global__ void k_coef_slots(const double* __restrict__ C, int Rsel, int N, int KP, int Rpad, double* __restrict__ Cs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Rpad * KP) return;
  const int r = (int)(e / KP), k = (int)(e - (int64_t)r * KP);
  Cs[(int64_t)r * KP + k_slot(k)] = (r < Rsel && k < N) ? C[(int64_t)r * N + k] : 0.0;
}

__global__ void k_fill_nan(double* __restrict__ out, int64_t n) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    for (; i + 1 < n; i += stride) *reinterpret_cast<double2*>(out + i) = make_double2(nan, nan);
    if (i < n) out[i] = nan;
  } else {
    for (; i < n; i += stride) { out[i] = nan; if (i + 1 < n) out[i + 1] = nan; }
  }
}

// out[r][idx[j]] = Arows[j] . Cs[r].  Persistent CTAs (one per SM: the A tile takes 150 KB of shared memory): the work
// units (128-point tile, 32-record chunk), tile-major, are split evenly into one contiguous range per CTA, so the
// number of in-hull points -- known only on the device -- never leaves a partial last wave, and a CTA reloads its
// A tile only when its range crosses into the next tile.
__global__ void __launch_bounds__(kGemmThreads)
k_est_gemm(const double* __restrict__ Arows, const int32_t* __restrict__ idx, const int32_t* __restrict__ count,
           const double* __restrict__ Cs, int Rsel, int Rpad, int KP, int LD, int64_t npts, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                               // kTileP x LD
  double* sC = smem + (size_t)kTileP * LD;         // 2 x kRC x LD
  __shared__ int32_t s_idx[kTileP];
  const int nin = *count;
  const int nchunk = Rpad / kRC;
  const int64_t total = (int64_t)((nin + kTileP - 1) / kTileP) * nchunk;
  const int64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
  if (u0 >= u1) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int pg = warp & 3, rh = warp >> 2;         // 32-point group, 16-record half of the chunk
  const int kp2 = KP >> 1;                         // 16-byte pieces per row
  auto load_chunk = [&](int chunk, int buf) {
    double* dst = sC + (size_t)buf * kRC * LD;
    const int r0 = chunk * kRC;
    for (int e = tid; e < kRC * kp2; e += kGemmThreads) {
      const int rr = e / kp2, q = e - rr * kp2;
      cpa16(dst + (size_t)rr * LD + 2 * q, Cs + (int64_t)(r0 + rr) * KP + 2 * q);      // Cs is padded to Rpad rows
    }
    cpa_commit();
  };
  int cur_tile = -1;
  load_chunk((int)(u0 % nchunk), 0);
  for (int64_t u = u0; u < u1; ++u) {
    const int tile = (int)(u / nchunk), ch = (int)(u - (int64_t)tile * nchunk);
    const int i = (int)(u - u0);
    if (tile != cur_tile) {                        // (every warp is past the previous tile: barrier at the loop's end)
      const int64_t j0 = (int64_t)tile * kTileP;
      const int rows = (nin - j0 < kTileP) ? (int)(nin - j0) : kTileP;
      if (tid < kTileP) s_idx[tid] = (tid < rows) ? idx[j0 + tid] : -1;
      // A tile: Arows is basis-major (k-slot x compacted point); two points per 16-byte load, transposed into the
      // point-major rows the fragments read.  Columns past `rows` hold whatever the workspace holds: their products
      // only reach accumulator rows that are never stored (s_idx = -1).
      const int64_t ldj = (npts + 1) & ~(int64_t)1;
#pragma unroll 4
      for (int e = tid; e < KP * (kTileP / 2); e += kGemmThreads) {
        const int k = e / (kTileP / 2), pp = e - k * (kTileP / 2);
        double2 v = make_double2(0.0, 0.0);
        if (j0 + 2 * pp + 1 < ldj) v = __ldg(reinterpret_cast<const double2*>(Arows + (int64_t)k * ldj + j0) + pp);
        sA[(size_t)(2 * pp) * LD + k] = (2 * pp < rows) ? v.x : 0.0;
        sA[(size_t)(2 * pp + 1) * LD + k] = (2 * pp + 1 < rows) ? v.y : 0.0;
      }
      cpa_commit();
      cur_tile = tile;
    }
    if (u + 1 < u1) { load_chunk((int)((u + 1) % nchunk), (i + 1) & 1); cpa_wait<1>(); } else cpa_wait<0>();
    __syncthreads();
    const double* Cb = sC + (size_t)(i & 1) * kRC * LD;
    double acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.0;
    for (int ks = 0; ks < KP / 16; ++ks) {
      const int kb = 16 * ks + 4 * t;
      double a[2][8];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double* q0 = sA + (size_t)(32 * pg + 16 * mt + g) * LD + kb;
        const double* q1 = q0 + 8 * LD;
        const double2 x01 = *reinterpret_cast<const double2*>(q0), x23 = *reinterpret_cast<const double2*>(q0 + 2);
        const double2 z01 = *reinterpret_cast<const double2*>(q1), z23 = *reinterpret_cast<const double2*>(q1 + 2);
        a[mt][0] = x01.x; a[mt][2] = x01.y; a[mt][4] = x23.x; a[mt][6] = x23.y;
        a[mt][1] = z01.x; a[mt][3] = z01.y; a[mt][5] = z23.x; a[mt][7] = z23.y;
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const double* pb = Cb + (size_t)(16 * rh + 8 * nt + g) * LD + kb;
        const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
        dmma_16x8x16(acc[0][nt], a[0], b);
        dmma_16x8x16(acc[1][nt], a[1], b);
      }
    }
    // c0 (point g, record 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int pl = 32 * pg + 16 * mt + g + ((v & 2) ? 8 : 0);
          const int r = ch * kRC + 16 * rh + 8 * nt + 2 * t + (v & 1);
          const int32_t p = s_idx[pl];
          if (p >= 0 && r < Rsel) out[(int64_t)r * npts + p] = acc[mt][nt][v];
        }
    __syncthreads();
  }
}

inline unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// One side stream per device for the NaN fill of the output tile: pure HBM writes that run beside the hull test and
// the basis rows (FP64 ALU work) instead of in front of the GEMM.
struct Side { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
Side g_side[64];
std::mutex g_side_mu;
int side_for_device(Side** out) {
  int dev = 0;
  VI_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { vi_set_error("device index %d out of range", dev); return VI_EINVAL; }
  std::lock_guard<std::mutex> lk(g_side_mu);
  Side& sd = g_side[dev];
  if (!sd.stream) {
    VI_CUDA(cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking));
    VI_CUDA(cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming));
    VI_CUDA(cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming));
  }
  *out = &sd;
  return VI_OK;
}

}  // namespace

// first half of vi_estimate_*_many (called before the hull test is launched): out <- NaN on the side stream
int vi_estimate_fill_begin(double* out, int32_t Rsel, int64_t npts, cudaStream_t s) {
  Side* sd = nullptr;
  if (int rc = side_for_device(&sd)) return rc;
  VI_CUDA(cudaEventRecord(sd->fork, s));                 // out may still be read by earlier work on s
  VI_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
  const int64_t pairs = ((int64_t)Rsel * npts + 1) / 2;
  // a few CTAs per SM: enough stores in flight for the HBM writes, most thread slots left to the basis rows
  static const int per_sm = getenv("VI_FILL_CTAS") ? atoi(getenv("VI_FILL_CTAS")) : 4;
  const unsigned cap = (unsigned)(per_sm > 0 ? per_sm : 4) * (unsigned)vi_sm_count();
  const unsigned grid = blocks(pairs, 256) < cap ? blocks(pairs, 256) : cap;
  VI_KERNEL(VI_K_ESTIMATE, sd->stream, k_fill_nan<<<grid, 256, 0, sd->stream>>>(out, (int64_t)Rsel * npts));
  VI_CUDA(cudaEventRecord(sd->join, sd->stream));
  return VI_OK;
}

// workspace layout (shared with basis.cu): [count: 256 B][idx: int32 x npts][Arows: KP x ldj, ldj = even(npts)][Cs: Rpad x KP]
extern "C" int vi_estimate_workspace_bytes(int64_t npts, int32_t N, int32_t Rsel, int64_t* bytes) {
  VI_REQUIRE(bytes != nullptr && npts >= 0 && N >= 1 && Rsel >= 1, "bad arguments");
  const int64_t KP = (N + 15) / 16 * 16, Rpad = (Rsel + kRC - 1) / kRC * kRC;
  *bytes = 256 + vi_align_up(npts * 4, 256) + (npts + 2) * KP * 8 + Rpad * KP * 8 + 256;
  return VI_OK;
}

// last part of vi_estimate_*_many: idx / count / Arows already filled by basis.cu, vi_estimate_fill_begin called
int vi_estimate_gemm_launch(const int32_t* count, const int32_t* idx, const double* Arows, double* Cs, const double* C,
                            int32_t Rsel, int32_t N, int64_t npts, double* out, cudaStream_t s) {
  const int KP = (N + 15) / 16 * 16, Rpad = (Rsel + kRC - 1) / kRC * kRC;
  int LD = KP + 2;
  while (LD % 16 != 2) ++LD;
  const size_t smem = ((size_t)kTileP * LD + 2 * (size_t)kRC * LD) * sizeof(double);
  if (smem > 227 * 1024 - 1024) { vi_set_error("nbasis %d too large for the Estimate GEMM tile", N); return VI_EUNSUPPORTED; }
  VI_KERNEL(VI_K_ESTIMATE, s, k_coef_slots<<<blocks((int64_t)Rpad * KP, 256), 256, 0, s>>>(C, Rsel, N, KP, Rpad, Cs));
  {
    Side* sd = nullptr;
    if (int rc = side_for_device(&sd)) return rc;
    VI_CUDA(cudaStreamWaitEvent(s, sd->join, 0));        // the NaN fill of vi_estimate_fill_begin is complete
  }
  VI_CUDA(cudaFuncSetAttribute(k_est_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VI_KERNEL(VI_K_EST_GEMM, s, k_est_gemm<<<(unsigned)vi_sm_count(), kGemmThreads, smem, s>>>(Arows, idx, count, Cs, Rsel, Rpad, KP, LD, npts, out));
  return VI_OK;
}
