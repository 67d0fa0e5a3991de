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

// Launch at the LOWEST stream priority of the device, whatever the priority of `st`: for the two large, issue bound
// rasterizer kernels.  A caller that issues independent views on high-priority streams (bench.py, distributed.run_views)
// thereby lets the CTAs of every other kernel of a view — sorts, scans, projection, SH: memory / latency bound, few issue
// slots — go first whenever an SM has room, so that they run UNDER another view's rasterizer instead of queueing behind
// its 11 000 CTAs.  On default-priority streams the attribute changes nothing.  The priority is a kernel node attribute
// under graph capture.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_background(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                            cudaStream_t st, Args&&... args) {
  static int least = 0, greatest = 0, have = 0;
  if (!have) { cudaDeviceGetStreamPriorityRange(&least, &greatest); have = 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributePriority;
  attr[0].val.priority = least;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
