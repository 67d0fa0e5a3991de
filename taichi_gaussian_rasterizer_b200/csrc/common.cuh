// common.cuh — shared helpers for the sm_100a kernels of libgsplat_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsplat_b200.h"

namespace gs {

void set_error(const char* fmt, ...);

#define GS_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      gs::set_error(__VA_ARGS__);      \
      return GS_ERR_INVALID;           \
    }                                  \
  } while (0)

#define GS_UNSUPPORTED(...)            \
  do {                                 \
    gs::set_error(__VA_ARGS__);        \
    return GS_ERR_UNSUPPORTED;         \
  } while (0)

#define GS_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      gs::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__,  \
                    __LINE__);                                                            \
      return GS_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define GS_LAUNCH_CHECK()                                                                 \
  do {                                                                                    \
    cudaError_t err__ = cudaGetLastError();                                               \
    if (err__ != cudaSuccess) {                                                           \
      gs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(err__),        \
                    __FILE__, __LINE__);                                                  \
      return GS_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

}  // namespace gs
