// K1 (design matrix) and K4 (Estimate) kernels for both model plug-ins.
//
//   vi_basis_sphharmlag   <- models/sphharmlag.py:118-145 (+ transform_coord :324-359)
//   vi_grad_basis_sphharmlag <- models/sphharmlag.py:148-184
//   vi_basis_radbasfun    <- models/radbasfun.py:83-112
//   vi_estimate_*         <- estimate.py:113-121 (basis . C, NaN outside the convex hull;
//                            hull test = facet half-spaces of the saved hull, estimate.py:153-178)
//
// The per-point mathematics lives in vi_math.h (shared with the CPU test harness); this
// file is compiled with -fmad=false so the operation order written there is what runs.
// One thread owns one query point: the special-function work (Legendre series + degree
// recurrence) is FP64-ALU bound and embarrassingly parallel, memory traffic is 24 B in and
// 8 B x (#records) out per point.
#include "common.cuh"
#include "vi_math.h"

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
k_basis_shl(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
            int64_t npts, const __grid_constant__ vi_shl_params P, double* __restrict__ A, double* __restrict__ At) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int N = P.maxk * P.maxl * P.maxl;
  double* row = A ? A + p * N : nullptr;
  double* col = At ? At + p : nullptr;
  vi_shl_row(P, lat[p], lon[p], alt[p], [&](int n, double v) {
    if (row) row[n] = v;
    if (col) col[(int64_t)n * npts] = v;
  });
}

__global__ void __launch_bounds__(kThreads)
k_basis_rbf(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
            int64_t npts, const double* __restrict__ centers, int N, double eps,
            double* __restrict__ A, double* __restrict__ At) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  double x, y, z;
  vi_geodetic2ecef(lat[p], lon[p], alt[p], &x, &y, &z);
  for (int n = 0; n < N; ++n) {
    double v = vi_rbf_value(x, y, z, centers[3 * n], centers[3 * n + 1], centers[3 * n + 2], eps);
    if (A) A[p * N + n] = v;
    if (At) At[(int64_t)n * npts + p] = v;
  }
}

// inside <=> n_f . x + d_f <= 0 for every facet f of ConvexHull(hull_vert)
__device__ __forceinline__ bool inside_hull(const double* __restrict__ eq, int F, double x, double y, double z) {
  bool in = true;
  for (int f = 0; f < F; ++f) {
    double s = eq[4 * f] * x + eq[4 * f + 1] * y + eq[4 * f + 2] * z + eq[4 * f + 3];
    in = in && (s <= 0.0);
  }
  return in;
}

// Few records (Rsel <= RT per pass): accumulate the dot products in registers while the basis
// values are produced; nothing but lat/lon/alt in and the results out touches memory.
template <int RT>
__global__ void __launch_bounds__(kThreads)
k_est_shl_reg(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
              int64_t npts, const __grid_constant__ vi_shl_params P, const double* __restrict__ C, int Rsel,
              const double* __restrict__ eq, int F, double* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int N = P.maxk * P.maxl * P.maxl;
  double la = lat[p], lo = lon[p], al = alt[p];
  bool in = true;
  if (F > 0) {
    double x, y, z;
    vi_geodetic2ecef(la, lo, al, &x, &y, &z);
    in = inside_hull(eq, F, x, y, z);
  }
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int r0 = 0; r0 < Rsel; r0 += RT) {
    double acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0;
    if (in) {
      vi_shl_row(P, la, lo, al, [&](int n, double v) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (r0 + r < Rsel) acc[r] = acc[r] + v * C[(int64_t)(r0 + r) * N + n];
      });
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (r0 + r < Rsel) out[(int64_t)(r0 + r) * npts + p] = in ? acc[r] : nan;
  }
}

// Many records: the CTA's 128 basis rows are staged in shared memory (transposed, n-major) once,
// then contracted against every record's coefficients (C is read through L1 as broadcasts).
// NOTE(order): the reference's einsum sums n = 0..N-1 sequentially; so does this loop.
constexpr int kEstRT = 16;
__global__ void __launch_bounds__(kThreads)
k_est_shl_tile(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
               int64_t npts, const __grid_constant__ vi_shl_params P, const double* __restrict__ C, int Rsel,
               const double* __restrict__ eq, int F, double* __restrict__ out) {
  extern __shared__ double sA[];   // N x kThreads
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P.maxk * P.maxl * P.maxl;
  const int t = threadIdx.x;
  bool in = false;
  if (p < npts) {
    double la = lat[p], lo = lon[p], al = alt[p];
    in = true;
    if (F > 0) {
      double x, y, z;
      vi_geodetic2ecef(la, lo, al, &x, &y, &z);
      in = inside_hull(eq, F, x, y, z);
    }
    if (in) vi_shl_row(P, la, lo, al, [&](int n, double v) { sA[n * kThreads + t] = v; });
  }
  if (p >= npts) return;   // no barrier needed: each thread only reads its own column
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int r0 = 0; r0 < Rsel; r0 += kEstRT) {
    double acc[kEstRT];
#pragma unroll
    for (int r = 0; r < kEstRT; ++r) acc[r] = 0.0;
    if (in) {
      for (int n = 0; n < N; ++n) {
        double a = sA[n * kThreads + t];
#pragma unroll
        for (int r = 0; r < kEstRT; ++r)
          if (r0 + r < Rsel) acc[r] = acc[r] + a * __ldg(C + (int64_t)(r0 + r) * N + n);
      }
    }
#pragma unroll
    for (int r = 0; r < kEstRT; ++r)
      if (r0 + r < Rsel) out[(int64_t)(r0 + r) * npts + p] = in ? acc[r] : nan;
  }
}

// Many records on the FP64 tensor cores: out[r][p] = sum_n basis(p)[n] C[r][n] is a GEMM
// [points x N] . [N x records].  The CTA evaluates the basis rows of its 128 points ONCE into shared
// memory, then streams the coefficient vectors through in chunks of 32 records (double buffered) and
// contracts with mma.sync.m16n8k16.f64.  Both operands are stored with the k index permuted inside each
// group of 16 (k -> 4 (k % 4) + k / 4) and a row stride = 2 (mod 16) doubles, so that every fragment is
// two conflict-free 128-bit shared loads (same scheme as k_ne_dmma3).
constexpr int kMmaRC = 32;          // records per chunk

__device__ __forceinline__ int k_slot(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }

__device__ __forceinline__ void est_dmma(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(kThreads)
k_est_shl_mma(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
              int64_t npts, const __grid_constant__ vi_shl_params P, const double* __restrict__ C, int Rsel,
              const double* __restrict__ eq, int F, double* __restrict__ out, int KP, int LD) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                              // kThreads x LD   basis rows (slot order)
  double* sC = smem + (size_t)kThreads * LD;      // 2 x kMmaRC x LD coefficient chunks (slot order)
  __shared__ unsigned char s_in[kThreads];
  const int N = P.maxk * P.maxl * P.maxl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t p0 = (int64_t)blockIdx.x * kThreads;
  const int64_t p = p0 + tid;
  // ---- phase 1: one thread per point evaluates its basis row into shared memory ----------------------
  {
    double* row = sA + (size_t)tid * LD;
    for (int k = N; k < KP; ++k) row[k_slot(k)] = 0.0;      // K padding
    bool in = false;
    if (p < npts) {
      const double la = lat[p], lo = lon[p], al = alt[p];
      in = true;
      if (F > 0) {
        double x, y, z;
        vi_geodetic2ecef(la, lo, al, &x, &y, &z);
        in = inside_hull(eq, F, x, y, z);
      }
      if (in) vi_shl_row(P, la, lo, al, [&](int n, double v) { row[k_slot(n)] = v; });
    }
    if (!in) for (int k = 0; k < N; ++k) row[k_slot(k)] = 0.0;
    s_in[tid] = in ? 1 : 0;
  }
  auto load_chunk = [&](int chunk, int buf) {
    double* dst = sC + (size_t)buf * kMmaRC * LD;
    const int r0 = chunk * kMmaRC;
    for (int e = tid; e < kMmaRC * KP; e += kThreads) {
      const int rr = e / KP, k = e - rr * KP;
      double v = 0.0;
      if (r0 + rr < Rsel && k < N) v = C[(int64_t)(r0 + rr) * N + k];
      dst[rr * LD + k_slot(k)] = v;
    }
  };
  const int nchunk = (Rsel + kMmaRC - 1) / kMmaRC;
  load_chunk(0, 0);
  __syncthreads();
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  // warp w owns points 32 w .. 32 w + 31 = two m16 tiles; four n8 tiles cover the 32 records of a chunk
  for (int ch = 0; ch < nchunk; ++ch) {
    if (ch + 1 < nchunk) load_chunk(ch + 1, (ch + 1) & 1);
    const double* Cb = sC + (size_t)(ch & 1) * kMmaRC * LD;
    double acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.0;
    for (int ks = 0; ks < KP / 16; ++ks) {
      const int kb = 16 * ks + 4 * t;
      double a[2][8];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double* q0 = sA + (size_t)(32 * warp + 16 * mt + g) * LD + kb;
        const double* q1 = q0 + 8 * LD;
        const double2 x01 = *reinterpret_cast<const double2*>(q0), x23 = *reinterpret_cast<const double2*>(q0 + 2);
        const double2 z01 = *reinterpret_cast<const double2*>(q1), z23 = *reinterpret_cast<const double2*>(q1 + 2);
        a[mt][0] = x01.x; a[mt][2] = x01.y; a[mt][4] = x23.x; a[mt][6] = x23.y;
        a[mt][1] = z01.x; a[mt][3] = z01.y; a[mt][5] = z23.x; a[mt][7] = z23.y;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double* pb = Cb + (size_t)(8 * nt + g) * LD + kb;
        const double2 b01 = *reinterpret_cast<const double2*>(pb), b23 = *reinterpret_cast<const double2*>(pb + 2);
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
        est_dmma(acc[0][nt], a[0], b);
        est_dmma(acc[1][nt], a[1], b);
      }
    }
    // c0 (point g, record 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int pl = 32 * warp + 16 * mt + g + ((v & 2) ? 8 : 0);
          const int r = ch * kMmaRC + 8 * nt + 2 * t + (v & 1);
          if (p0 + pl < npts && r < Rsel) out[(int64_t)r * npts + p0 + pl] = s_in[pl] ? acc[mt][nt][v] : nan;
        }
    __syncthreads();
  }
}

template <int RT>
__global__ void __launch_bounds__(kThreads)
k_est_rbf(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
          int64_t npts, const double* __restrict__ centers, int N, double eps, const double* __restrict__ C,
          int Rsel, const double* __restrict__ eq, int F, double* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  double x, y, z;
  vi_geodetic2ecef(lat[p], lon[p], alt[p], &x, &y, &z);
  bool in = (F > 0) ? inside_hull(eq, F, x, y, z) : true;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int r0 = 0; r0 < Rsel; r0 += RT) {
    double acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0;
    if (in) {
      for (int n = 0; n < N; ++n) {
        double v = vi_rbf_value(x, y, z, centers[3 * n], centers[3 * n + 1], centers[3 * n + 2], eps);
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (r0 + r < Rsel) acc[r] = acc[r] + v * __ldg(C + (int64_t)(r0 + r) * N + n);
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (r0 + r < Rsel) out[(int64_t)(r0 + r) * npts + p] = in ? acc[r] : nan;
  }
}

// ---- many-record Estimate, first half (estimate_gemm.cu has the second): hull compaction + basis rows ----------
__device__ __forceinline__ int k_slot16(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }

// idx[0 .. *count) = points inside the hull (order arbitrary: every point's result is independent of it)
__global__ void __launch_bounds__(256)
k_hull_compact(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
               int64_t npts, const double* __restrict__ eq, int F, int32_t* __restrict__ idx, int32_t* __restrict__ count) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  if (p < npts) {
    in = true;
    if (F > 0) {
      double x, y, z;
      vi_geodetic2ecef(lat[p], lon[p], alt[p], &x, &y, &z);
      in = inside_hull(eq, F, x, y, z);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, in);
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (in) idx[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)p;
}

// basis row of compacted point j, stored TRANSPOSED: value n of point j at Arows[slot(n) * ldj + j] (slot order:
// k -> 4 (k % 4) + k / 4 inside each group of 16; zero padded to KP).  Consecutive threads write consecutive addresses:
// with the point-major layout every store of a warp touched 32 sectors, and the kernel could not share the memory
// system with the NaN fill that runs beside it.
__global__ void __launch_bounds__(kThreads)
k_rows_shl_idx(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
               const int32_t* __restrict__ idx, const int32_t* __restrict__ count, const __grid_constant__ vi_shl_params P,
               int KP, int64_t ldj, double* __restrict__ Arows) {
  // one thread per (compacted point, degree l = blockIdx.y): the in-hull points alone are too few threads to hide the
  // latency of the Legendre series (490 per SM on a 2^17-point tile); the degrees are independent
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *count) return;
  const int N = P.maxk * P.maxl * P.maxl;
  const int32_t p = idx[j];
  const int l = blockIdx.y;
  double* col = Arows + j;
  if (l == 0)
    for (int k = N; k < KP; ++k) col[k_slot16(k) * ldj] = 0.0;
  vi_shl_row(P, lat[p], lon[p], alt[p], [&](int n, double v) { col[k_slot16(n) * ldj] = v; }, l, l + 1);
}

__global__ void __launch_bounds__(kThreads)
k_rows_rbf_idx(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
               const int32_t* __restrict__ idx, const int32_t* __restrict__ count, const double* __restrict__ centers,
               int N, double eps, int KP, int64_t ldj, double* __restrict__ Arows) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *count) return;
  const int32_t p = idx[j];
  double x, y, z;
  vi_geodetic2ecef(lat[p], lon[p], alt[p], &x, &y, &z);
  double* col = Arows + j;
  for (int k = N; k < KP; ++k) col[k_slot16(k) * ldj] = 0.0;
  for (int n = 0; n < N; ++n)
    col[k_slot16(n) * ldj] = vi_rbf_value(x, y, z, centers[3 * n], centers[3 * n + 1], centers[3 * n + 2], eps);
}

int check_shl(const vi_shl_params* P) {
  VI_REQUIRE(P != nullptr, "params is NULL");
  VI_REQUIRE(P->maxk >= 1 && P->maxk <= VI_MAXK_MAX && P->maxl >= 1 && P->maxl <= VI_MAXL_MAX,
             "MAXK/MAXL out of range (1..%d / 1..%d)", VI_MAXK_MAX, VI_MAXL_MAX);
  return VI_OK;
}

// Gradient of the basis, out[p][comp][n] (npts x 3 x N, the shape the reference returns): one thread per point.
__global__ void __launch_bounds__(kThreads)
k_grad_shl(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ alt,
           int64_t npts, const __grid_constant__ vi_shl_params P, double* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const int N = P.maxk * P.maxl * P.maxl;
  double* o = out + p * 3 * (int64_t)N;
  vi_shl_grad_row(P, lat[p], lon[p], alt[p], [&](int n, double gz, double gt, double gp) {
    o[n] = gz; o[N + n] = gt; o[2 * N + n] = gp;
  });
}

}  // namespace

extern "C" int vi_grad_basis_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                                        const vi_shl_params* params, double* out, void* stream) {
  if (int rc = check_shl(params)) return rc;
  VI_REQUIRE(npts >= 0 && (out != nullptr || npts == 0), "bad arguments");
  if (npts == 0) return VI_OK;
  unsigned grid = (unsigned)((npts + kThreads - 1) / kThreads);
  VI_KERNEL(VI_K_BASIS, vi_stream(stream), k_grad_shl<<<grid, kThreads, 0, vi_stream(stream)>>>(lat, lon, alt, npts, *params, out));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_basis_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                                   const vi_shl_params* params, double* A, double* At, void* stream) {
  if (int rc = check_shl(params)) return rc;
  VI_REQUIRE(npts >= 0, "npts < 0");
  if (npts == 0) return VI_OK;
  unsigned grid = (unsigned)((npts + kThreads - 1) / kThreads);
  VI_KERNEL(VI_K_BASIS, vi_stream(stream), k_basis_shl<<<grid, kThreads, 0, vi_stream(stream)>>>(lat, lon, alt, npts, *params, A, At));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_basis_radbasfun(const double* lat, const double* lon, const double* alt, int64_t npts,
                                  const double* centers, int32_t N, double eps, double* A, double* At, void* stream) {
  VI_REQUIRE(npts >= 0 && N >= 1 && centers != nullptr, "bad arguments");
  if (npts == 0) return VI_OK;
  unsigned grid = (unsigned)((npts + kThreads - 1) / kThreads);
  VI_KERNEL(VI_K_BASIS, vi_stream(stream), k_basis_rbf<<<grid, kThreads, 0, vi_stream(stream)>>>(lat, lon, alt, npts, centers, N, eps, A, At));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_estimate_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                                      const vi_shl_params* params, const double* C, int32_t Rsel,
                                      const double* hull_eq, int32_t F, double* out, void* stream) {
  if (int rc = check_shl(params)) return rc;
  VI_REQUIRE(npts >= 0 && Rsel >= 1 && C != nullptr && out != nullptr, "bad arguments");
  if (npts == 0) return VI_OK;
  if (hull_eq == nullptr) F = 0;
  unsigned grid = (unsigned)((npts + kThreads - 1) / kThreads);
  const int N = params->maxk * params->maxl * params->maxl;
  cudaStream_t s = vi_stream(stream);
  size_t smem = (size_t)N * kThreads * sizeof(double);
  if (Rsel == 1) {
    VI_KERNEL(VI_K_ESTIMATE, s, k_est_shl_reg<1><<<grid, kThreads, 0, s>>>(lat, lon, alt, npts, *params, C, Rsel, hull_eq, F, out));
  } else if (Rsel <= 8 || smem > 220 * 1024) {
    VI_KERNEL(VI_K_ESTIMATE, s, k_est_shl_reg<8><<<grid, kThreads, 0, s>>>(lat, lon, alt, npts, *params, C, Rsel, hull_eq, F, out));
  } else {
    const int KP = (N + 15) / 16 * 16;
    int LD = KP + 2;
    while (LD % 16 != 2) ++LD;
    const size_t smem_mma = ((size_t)kThreads * LD + 2 * (size_t)kMmaRC * LD) * sizeof(double);
    if (Rsel >= 16 && smem_mma <= 227 * 1024 - 256) {
      VI_CUDA(cudaFuncSetAttribute(k_est_shl_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
      VI_KERNEL(VI_K_ESTIMATE, s, k_est_shl_mma<<<grid, kThreads, smem_mma, s>>>(lat, lon, alt, npts, *params, C, Rsel, hull_eq, F, out, KP, LD));
    } else {
      VI_CUDA(cudaFuncSetAttribute(k_est_shl_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VI_KERNEL(VI_K_ESTIMATE, s, k_est_shl_tile<<<grid, kThreads, smem, s>>>(lat, lon, alt, npts, *params, C, Rsel, hull_eq, F, out));
    }
  }
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_estimate_radbasfun(const double* lat, const double* lon, const double* alt, int64_t npts,
                                     const double* centers, int32_t N, double eps, const double* C, int32_t Rsel,
                                     const double* hull_eq, int32_t F, double* out, void* stream) {
  VI_REQUIRE(npts >= 0 && Rsel >= 1 && N >= 1 && C != nullptr && out != nullptr && centers != nullptr, "bad arguments");
  if (npts == 0) return VI_OK;
  if (hull_eq == nullptr) F = 0;
  unsigned grid = (unsigned)((npts + kThreads - 1) / kThreads);
  cudaStream_t s = vi_stream(stream);
  if (Rsel == 1)
    VI_KERNEL(VI_K_ESTIMATE, s, k_est_rbf<1><<<grid, kThreads, 0, s>>>(lat, lon, alt, npts, centers, N, eps, C, Rsel, hull_eq, F, out));
  else
    VI_KERNEL(VI_K_ESTIMATE, s, k_est_rbf<8><<<grid, kThreads, 0, s>>>(lat, lon, alt, npts, centers, N, eps, C, Rsel, hull_eq, F, out));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

// estimate_gemm.cu: NaN fill of the output on a side stream (first), coefficient slots + GEMM (last)
int vi_estimate_fill_begin(double* out, int32_t Rsel, int64_t npts, cudaStream_t s);
int vi_estimate_gemm_launch(const int32_t* count, const int32_t* idx, const double* Arows, double* Cs, const double* C,
                            int32_t Rsel, int32_t N, int64_t npts, double* out, cudaStream_t s);

namespace {
struct EstWs { int32_t* count; int32_t* idx; double* Arows; double* Cs; };
int est_carve(void* workspace, int64_t workspace_bytes, int64_t npts, int32_t N, int32_t Rsel, EstWs* w) {
  int64_t need = 0;
  if (int rc = vi_estimate_workspace_bytes(npts, N, Rsel, &need)) return rc;
  if (workspace == nullptr || workspace_bytes < need) {
    vi_set_error("estimate workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return VI_EWORKSPACE;
  }
  VI_REQUIRE(npts < ((int64_t)1 << 31), "at most 2^31 - 1 points per call (tile the grid)");
  const int64_t KP = (N + 15) / 16 * 16;
  char* b = reinterpret_cast<char*>(workspace);
  w->count = reinterpret_cast<int32_t*>(b); b += 256;
  w->idx = reinterpret_cast<int32_t*>(b); b += vi_align_up(npts * 4, 256);
  w->Arows = reinterpret_cast<double*>(b); b += (npts + 2) * KP * 8;      // KP x ldj, ldj = npts rounded up to even
  w->Cs = reinterpret_cast<double*>(b);
  return VI_OK;
}
}  // namespace

extern "C" int vi_estimate_sphharmlag_many(const double* lat, const double* lon, const double* alt, int64_t npts,
                                           const vi_shl_params* params, const double* C, int32_t Rsel,
                                           const double* hull_eq, int32_t F, double* out,
                                           void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_shl(params)) return rc;
  VI_REQUIRE(npts >= 0 && Rsel >= 1 && C != nullptr && out != nullptr, "bad arguments");
  if (npts == 0) return VI_OK;
  if (hull_eq == nullptr) F = 0;
  const int N = params->maxk * params->maxl * params->maxl;
  const int KP = (N + 15) / 16 * 16;
  EstWs w;
  if (int rc = est_carve(workspace, workspace_bytes, npts, N, Rsel, &w)) return rc;
  cudaStream_t s = vi_stream(stream);
  if (int rc = vi_estimate_fill_begin(out, Rsel, npts, s)) return rc;
  VI_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t), s));
  VI_KERNEL(VI_K_ESTIMATE, s, k_hull_compact<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(lat, lon, alt, npts, hull_eq, F, w.idx, w.count));
  VI_KERNEL(VI_K_ESTIMATE, s, k_rows_shl_idx<<<dim3((unsigned)((npts + kThreads - 1) / kThreads), (unsigned)params->maxl), kThreads, 0, s>>>(lat, lon, alt, w.idx, w.count, *params, KP, (npts + 1) & ~(int64_t)1, w.Arows));
  return vi_estimate_gemm_launch(w.count, w.idx, w.Arows, w.Cs, C, Rsel, N, npts, out, s);
}

extern "C" int vi_estimate_radbasfun_many(const double* lat, const double* lon, const double* alt, int64_t npts,
                                          const double* centers, int32_t N, double eps, const double* C, int32_t Rsel,
                                          const double* hull_eq, int32_t F, double* out,
                                          void* workspace, int64_t workspace_bytes, void* stream) {
  VI_REQUIRE(npts >= 0 && Rsel >= 1 && N >= 1 && C != nullptr && out != nullptr && centers != nullptr, "bad arguments");
  if (npts == 0) return VI_OK;
  if (hull_eq == nullptr) F = 0;
  const int KP = (N + 15) / 16 * 16;
  EstWs w;
  if (int rc = est_carve(workspace, workspace_bytes, npts, N, Rsel, &w)) return rc;
  cudaStream_t s = vi_stream(stream);
  if (int rc = vi_estimate_fill_begin(out, Rsel, npts, s)) return rc;
  VI_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int32_t), s));
  VI_KERNEL(VI_K_ESTIMATE, s, k_hull_compact<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(lat, lon, alt, npts, hull_eq, F, w.idx, w.count));
  VI_KERNEL(VI_K_ESTIMATE, s, k_rows_rbf_idx<<<(unsigned)((npts + kThreads - 1) / kThreads), kThreads, 0, s>>>(lat, lon, alt, w.idx, w.count, centers, N, eps, KP, (npts + 1) & ~(int64_t)1, w.Arows));
  return vi_estimate_gemm_launch(w.count, w.idx, w.Arows, w.Cs, C, Rsel, N, npts, out, s);
}

extern "C" int vi_estimate_sphharmlag_host(const double* lat, const double* lon, const double* alt, int64_t npts,
                                           const vi_shl_params* params, const double* C, int32_t Rsel,
                                           const double* hull_eq, int32_t F, double* out) {
  if (int rc = check_shl(params)) return rc;
  VI_REQUIRE(npts >= 0 && Rsel >= 1, "bad arguments");
  if (npts == 0) return VI_OK;
  const int N = params->maxk * params->maxl * params->maxl;
  double *d_in = nullptr, *d_C = nullptr, *d_eq = nullptr, *d_out = nullptr;
  cudaStream_t s = nullptr;
  int rc = VI_OK;
  size_t nb = (size_t)npts * sizeof(double);
  if (hull_eq == nullptr) F = 0;
  // (no early returns between the allocations and the frees: a failed call must not leak device memory)
#define VI_TRY(x) do { if (rc == VI_OK) { cudaError_t _e = (x); if (_e != cudaSuccess) { vi_set_error("%s -> %s", #x, cudaGetErrorString(_e)); rc = VI_ECUDA; } } } while (0)
  VI_TRY(cudaMalloc(&d_in, 3 * nb));
  VI_TRY(cudaMalloc(&d_C, (size_t)Rsel * N * sizeof(double)));
  VI_TRY(cudaMalloc(&d_out, (size_t)Rsel * nb));
  if (F > 0) VI_TRY(cudaMalloc(&d_eq, (size_t)F * 4 * sizeof(double)));
  VI_TRY(cudaMemcpyAsync(d_in, lat, nb, cudaMemcpyHostToDevice, s));
  VI_TRY(cudaMemcpyAsync(d_in + npts, lon, nb, cudaMemcpyHostToDevice, s));
  VI_TRY(cudaMemcpyAsync(d_in + 2 * npts, alt, nb, cudaMemcpyHostToDevice, s));
  VI_TRY(cudaMemcpyAsync(d_C, C, (size_t)Rsel * N * sizeof(double), cudaMemcpyHostToDevice, s));
  if (F > 0) VI_TRY(cudaMemcpyAsync(d_eq, hull_eq, (size_t)F * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  if (rc == VI_OK) rc = vi_estimate_sphharmlag(d_in, d_in + npts, d_in + 2 * npts, npts, params, d_C, Rsel, d_eq, F, d_out, s);
  VI_TRY(cudaMemcpyAsync(out, d_out, (size_t)Rsel * nb, cudaMemcpyDeviceToHost, s));
  VI_TRY(cudaStreamSynchronize(s));
#undef VI_TRY
  if (d_in) cudaFree(d_in);
  if (d_C) cudaFree(d_C);
  if (d_out) cudaFree(d_out);
  if (d_eq) cudaFree(d_eq);
  return rc;
}
