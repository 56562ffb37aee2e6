// Shared host-side plumbing of the C-ABI translation units: error string, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include "volinterp_b200.h"

void vi_set_error(const char* fmt, ...);

#define VI_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      vi_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));    \
      return VI_ECUDA;                                                                       \
    }                                                                                        \
  } while (0)

#define VI_LAUNCH_CHECK()                                                                    \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      vi_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));\
      return VI_ECUDA;                                                                       \
    }                                                                                        \
  } while (0)

#define VI_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      vi_set_error(__VA_ARGS__);                   \
      return VI_EINVAL;                            \
    }                                              \
  } while (0)

static inline cudaStream_t vi_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t vi_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- launch accounting / optional per-kernel timing (diagnostics for bench.py) -----------------
// Every kernel launch of the library goes through VI_KERNEL(kind, stream, launch-expression): the
// launch counter of `kind` is always incremented; when vi_profile_enable(1) was called, the launch
// is bracketed by CUDA events on the launching stream and vi_profile_read() reports the summed
// device time per kind.  Not thread safe (one profiling client at a time).
int vi_sm_count();      // abi.cu

enum ViKind { VI_K_BASIS = 0, VI_K_NORMAL_EQ, VI_K_TRIDIAG, VI_K_TQL, VI_K_APPLY, VI_K_CHI2, VI_K_COV, VI_K_ESTIMATE, VI_K_MISC, VI_K_CHASE, VI_K_EST_GEMM, VI_K_COUNT };
void vi_prof_count_rotations(int64_t rotations, int64_t systems);
void vi_prof_launch_begin(int kind, cudaStream_t s);
void vi_prof_launch_end(int kind, cudaStream_t s);
#define VI_KERNEL(kind, stream, ...)          \
  do {                                        \
    vi_prof_launch_begin((kind), (stream));   \
    __VA_ARGS__;                              \
    vi_prof_launch_end((kind), (stream));     \
    VI_LAUNCH_CHECK();                        \
  } while (0)
