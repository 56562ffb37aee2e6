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
