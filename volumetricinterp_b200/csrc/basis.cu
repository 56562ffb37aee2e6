// K1 (design matrix) and K4 (Estimate) kernels for both model plug-ins.
//
//   vi_basis_sphharmlag   <- models/sphharmlag.py:118-145 (+ transform_coord :324-359)
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

int check_shl(const vi_shl_params* P) {
  VI_REQUIRE(P != nullptr, "params is NULL");
  VI_REQUIRE(P->maxk >= 1 && P->maxk <= VI_MAXK_MAX && P->maxl >= 1 && P->maxl <= VI_MAXL_MAX,
             "MAXK/MAXL out of range (1..%d / 1..%d)", VI_MAXK_MAX, VI_MAXL_MAX);
  return VI_OK;
}

}  // namespace

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
    VI_CUDA(cudaFuncSetAttribute(k_est_shl_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VI_KERNEL(VI_K_ESTIMATE, s, k_est_shl_tile<<<grid, kThreads, smem, s>>>(lat, lon, alt, npts, *params, C, Rsel, hull_eq, F, out));
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
  VI_CUDA(cudaMalloc(&d_in, 3 * nb));
  VI_CUDA(cudaMalloc(&d_C, (size_t)Rsel * N * sizeof(double)));
  VI_CUDA(cudaMalloc(&d_out, (size_t)Rsel * nb));
  if (F > 0) VI_CUDA(cudaMalloc(&d_eq, (size_t)F * 4 * sizeof(double)));
  VI_CUDA(cudaMemcpyAsync(d_in, lat, nb, cudaMemcpyHostToDevice, s));
  VI_CUDA(cudaMemcpyAsync(d_in + npts, lon, nb, cudaMemcpyHostToDevice, s));
  VI_CUDA(cudaMemcpyAsync(d_in + 2 * npts, alt, nb, cudaMemcpyHostToDevice, s));
  VI_CUDA(cudaMemcpyAsync(d_C, C, (size_t)Rsel * N * sizeof(double), cudaMemcpyHostToDevice, s));
  if (F > 0) VI_CUDA(cudaMemcpyAsync(d_eq, hull_eq, (size_t)F * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  rc = vi_estimate_sphharmlag(d_in, d_in + npts, d_in + 2 * npts, npts, params, d_C, Rsel, d_eq, F, d_out, s);
  if (rc == VI_OK) {
    VI_CUDA(cudaMemcpyAsync(out, d_out, (size_t)Rsel * nb, cudaMemcpyDeviceToHost, s));
    VI_CUDA(cudaStreamSynchronize(s));
  }
  cudaFree(d_in); cudaFree(d_C); cudaFree(d_out); if (d_eq) cudaFree(d_eq);
  return rc;
}
