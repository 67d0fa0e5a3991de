// Microbenchmark: issue / pipe throughput of FFMA vs the packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

constexpr int ITERS = 4096;
constexpr int CH = 8;

// MODE 0: 8 chains of FFMA.  1: 8 chains of FFMA2.  2: FFMA + 1 MUFU per 8.  3: FFMA2 + 1 MUFU per 8.
// 4: 4 FFMA2 + 4 FFMA interleaved.
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ x, float* y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float a = x[i & 1023], b = x[(i + 1) & 1023];
  float s[CH];
  uint64_t p[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { s[c] = a + c; p[c] = pk(a + c, b + c); }
  const uint64_t A2 = pk(a, b), B2 = pk(b, a);
  float m = a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (MODE == 0 || MODE == 2) s[c] = fma1(s[c], a, b);
      if (MODE == 1 || MODE == 3) p[c] = fma2(p[c], A2, B2);
      if (MODE == 4) { if (c & 1) s[c] = fma1(s[c], a, b); else p[c] = fma2(p[c], A2, B2); }
    }
    if (MODE == 2 || MODE == 3) m = ex2(m);
  }
  float r = m;
#pragma unroll
  for (int c = 0; c < CH; ++c) { float u, v; upk(p[c], u, v); r += s[c] + u + v; }
  y[i] = r;
}

template <int MODE>
void run(const char* name, const float* x, float* y, double fma_per_thread_iter) {
  const int blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(x, y);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(x, y);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double warp_instr = (double)blocks * threads / 32 * ITERS * CH;
  const double fmas = (double)blocks * threads * ITERS * fma_per_thread_iter;
  printf("{\"mode\": \"%s\", \"ms\": %.4f, \"warp_fp_instr_per_s\": %.4g, \"fma_lane_ops_per_s\": %.4g, \"tflops\": %.2f}\n",
         name, ms, warp_instr / (ms * 1e-3), fmas / (ms * 1e-3), 2 * fmas / (ms * 1e-3) / 1e12);
}

int main() {
  float *x, *y;
  cudaMalloc(&x, 1024 * 4); cudaMalloc(&y, 148 * 8 * 256 * 4);
  cudaMemset(x, 0, 1024 * 4);
  run<0>("ffma", x, y, CH);
  run<1>("ffma2", x, y, 2 * CH);
  run<2>("ffma+mufu", x, y, CH);
  run<3>("ffma2+mufu", x, y, 2 * CH);
  run<4>("ffma/ffma2 mix", x, y, 1.5 * CH);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
