// morton.cu — Morton (Z-order) codes of 3D points on a uniform grid, for spatial reordering of the gaussians.
//
// Replaces code_points32_kernel / code_points64_kernel of /root/reference/taichi_splatting/misc/morton_sort.py:93-111
// (Grid.grid_cell :51-54, spreads_bits32/64 :13-30, cell_code32/64 :69-88).  Byte / integer work bound by HBM:
// 12 B read + 4 or 8 B written per point.  The cell arithmetic is two IEEE f32 operations per axis, (p - lower) / inc,
// written with explicit round-to-nearest intrinsics so that the codes are bit-identical to the oracle's.
#include "common.cuh"

namespace gs {

__device__ __forceinline__ uint32_t spread_bits32(uint32_t x) {  // 10 bits -> every third bit
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__device__ __forceinline__ uint64_t spread_bits64(uint64_t x) {  // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// clamp((p - lower) / inc, 0, size - 1) truncated to an unsigned cell index; NaN goes to cell 0 like fmaxf(NaN, 0).
__device__ __forceinline__ uint32_t grid_cell(float p, float lower, float inc, float size_minus_1) {
  const float v = __fdiv_rn(__fsub_rn(p, lower), inc);
  return (uint32_t)fminf(fmaxf(v, 0.f), size_minus_1);
}

template <typename CodeT>
__global__ void __launch_bounds__(256)
morton_codes_kernel(int64_t n, const float* __restrict__ points, const float* __restrict__ lower,
                    const float* __restrict__ inc, float size_minus_1, CodeT* __restrict__ codes) {
  __shared__ float s_p[256 * 3];
  const float lx = lower[0], ly = lower[1], lz = lower[2], ix = inc[0], iy = inc[1], iz = inc[2];
  const int64_t base = (int64_t)blockIdx.x * 256;
  // coalesced read of the block's 256 xyz triples through shared memory
  for (int k = threadIdx.x; k < 256 * 3; k += 256) {
    const int64_t g = base * 3 + k;
    s_p[k] = g < n * 3 ? points[g] : 0.f;
  }
  __syncthreads();
  const int64_t i = base + threadIdx.x;
  if (i >= n) return;
  const uint32_t cx = grid_cell(s_p[3 * threadIdx.x], lx, ix, size_minus_1);
  const uint32_t cy = grid_cell(s_p[3 * threadIdx.x + 1], ly, iy, size_minus_1);
  const uint32_t cz = grid_cell(s_p[3 * threadIdx.x + 2], lz, iz, size_minus_1);
  if (sizeof(CodeT) == 8)
    codes[i] = (CodeT)(spread_bits64(cx) | (spread_bits64(cy) << 1) | (spread_bits64(cz) << 2));
  else
    codes[i] = (CodeT)(spread_bits32(cx) | (spread_bits32(cy) << 1) | (spread_bits32(cz) << 2));
}

}  // namespace gs

using namespace gs;

extern "C" int gs_morton_codes(int64_t n, const float* points, const float* lower, const float* inc,
                               int64_t grid_size, int32_t code_bits, void* codes, void* stream) {
  GS_CHECK_ARG(n >= 0, "gs_morton_codes: negative point count");
  GS_CHECK_ARG(code_bits == 32 || code_bits == 64, "gs_morton_codes: code_bits must be 32 or 64");
  GS_CHECK_ARG(grid_size >= 1 && grid_size <= (code_bits == 32 ? (1ll << 10) : (1ll << 21)),
               "gs_morton_codes: grid size %lld does not fit %d bit codes", (long long)grid_size, code_bits);
  if (n == 0) return GS_OK;
  GS_CHECK_ARG(points && lower && inc && codes, "gs_morton_codes: null tensor");
  const unsigned blocks = (unsigned)ceil_div(n, (int64_t)256);
  const float sm1 = (float)(grid_size - 1);
  if (code_bits == 64)
    morton_codes_kernel<uint64_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, points, lower, inc, sm1, (uint64_t*)codes);
  else
    morton_codes_kernel<uint32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, points, lower, inc, sm1, (uint32_t*)codes);
  GS_LAUNCH_CHECK();
  return GS_OK;
}
