// FP64 peak probes for the roofline denominators (MEASURED_PEAKS.json has HBM and bf16 only).
//   mode 0: independent DFMA chains (FP64 FMA pipe)
//   mode 1: mma.sync.m16n8k4 f64 (DMMA, the instruction K2 uses)
//   mode 2: mma.sync.m8n8k4  f64
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_probe_fma(int iters, double* out) {
  double a[8];
  const double x = 1.0 + 1e-9 * threadIdx.x, yv = 1e-9 * blockIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, yv);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_probe_dmma16(int iters, double* out) {
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
  const double a0 = 1e-3 * threadIdx.x, a1 = 2e-3, b0 = 1e-3 * (blockIdx.x + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(b0));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_probe_dmma8(int iters, double* out) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; }
  const double a0 = 1e-3 * threadIdx.x, b0 = 1e-3 * (blockIdx.x + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a0), "d"(b0));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace

extern "C" int vi_fp64_peak_probe(int32_t mode, int32_t iters, double* tflops, void* stream) {
  VI_REQUIRE(tflops != nullptr && iters > 0 && mode >= 0 && mode <= 2, "bad arguments");
  cudaStream_t s = vi_stream(stream);
  int dev = 0, sms = 0;
  VI_CUDA(cudaGetDevice(&dev));
  VI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d_out = nullptr;
  VI_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  VI_CUDA(cudaEventCreate(&e0));
  VI_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    VI_CUDA(cudaEventRecord(e0, s));
    if (mode == 0) k_probe_fma<<<grid, threads, 0, s>>>(iters, d_out);
    else if (mode == 1) k_probe_dmma16<<<grid, threads, 0, s>>>(iters, d_out);
    else k_probe_dmma8<<<grid, threads, 0, s>>>(iters, d_out);
    VI_CUDA(cudaEventRecord(e1, s));
    VI_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    VI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double flops;
    if (mode == 0) flops = 2.0 * 8 * (double)iters * threads * grid;
    else if (mode == 1) flops = 2.0 * 16 * 8 * 4 * 4 * (double)iters * (threads / 32) * grid;
    else flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * (threads / 32) * grid;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  *tflops = best;
  return VI_OK;
}
