// raster_fast.cuh — shared pieces of the fast (f32, alpha blending, tile 16) rasterizer kernels:
// packed per-gaussian records, cp.async staging, the lane-parallel ellipse / pixel-block cull test.
#pragma once

#include "raster.cuh"

namespace gs {

constexpr int kFastTile = 16;
constexpr int kFastTileArea = kFastTile * kFastTile;

// padded feature row length used by the fast kernels for a given F (0 = unsupported)
inline int fast_feature_pad(int F) { return F <= 4 ? 4 : F <= 8 ? 8 : F <= 16 ? 16 : F <= 36 ? 36 : F <= 64 ? 64 : 0; }

// Workspace layout (all 256 B aligned):
//   recF  V x 2 float4 : {mx, my, a1x, a1y} {a2x, a2y, log2(alpha), idx}       (forward records)
//   featP V x FP float  : features padded to FP
//   mask  K bytes        : per tile-list entry, which of the tile's eight 8x4 pixel blocks it can reach
//   recB  V x 2 float4 : {mx, my, ax, ay} {1/sx, 1/sy, alpha, idx}             (records of the wide backward, F > 7;
//                                                                               the narrow one walks recF)
struct FastLayout {
  size_t off_recF, off_feat, off_recB, off_mask, total;
  int FP;
};

inline FastLayout fast_layout(const GsRasterParams& p) {
  FastLayout L;
  L.FP = fast_feature_pad(p.num_features);
  const size_t V = (size_t)(p.num_points > 0 ? p.num_points : 1);
  size_t off = 0;
  L.off_recF = off; off += align_up(V * 32, 256);
  L.off_feat = off; off += align_up(V * (size_t)L.FP * 4, 256);
  L.off_recB = off; off += align_up(V * 32, 256);
  // one byte per tile-list entry: bit w = the entry can reach the 8x4 pixel block w of its tile (raster_cull_mask)
  L.off_mask = off; off += align_up((size_t)(p.num_overlaps > 0 ? p.num_overlaps : 1), 256);
  L.total = off;
  return L;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Logistic approximation of the normal CDF used by the antialiased pdf (taichi_lib/generic.py:340-352 S_sig):
// 1 / (1 + exp(-1.6 z - 0.07 z^3)).  aa_sig_grad also returns d/dz.  The pixel integral is a DIFFERENCE of two such
// terms (cancellation: 0.04 .. 0.3 of their size), so the 2 ulp of ex2.approx / rcp.approx the plain gaussian path
// lives with would leave the image 1.2e-5 from the oracle (measured); expf and the IEEE reciprocal bring it to 4e-6.
__device__ __forceinline__ float aa_sig(float z) {
  return __frcp_rn(1.f + expf(-z * fmaf(0.07f, z * z, 1.6f)));
}
__device__ __forceinline__ float aa_sig_grad(float z, float& ds_dz) {
  const float s = aa_sig(z);
  ds_dz = fmaf(0.21f, z * z, 1.6f) * s * (1.f - s);
  return s;
}

// exp(-0.5 (tx^2 + ty^2)) == exp2(-(c tx)^2 - (c ty)^2) with c = sqrt(0.5 log2(e))
constexpr float kSqrtHalfLog2e = 0.8493218002880191f;
constexpr float kHalfLog2e = 0.7213475204444817f;

// Conservative test: can gaussian (conic A of the scaled axes, exponent budget qmax) reach any pixel centre
// of the block [x0, x1] x [y0, y1]?  Minimises the quadratic form over the rectangle exactly (interior, or the
// clamped 1-D minimum on each edge facing the mean) and adds slack for rounding; the per pixel test stays exact.
__device__ __forceinline__ bool block_may_touch(float mx, float my, float a1x, float a1y, float a2x, float a2y,
                                                float qmax, float x0, float x1, float y0, float y1) {
  const float dx0 = x0 - mx, dx1 = x1 - mx, dy0 = y0 - my, dy1 = y1 - my;
  const float dxc = fminf(fmaxf(0.f, dx0), dx1);
  const float dyc = fminf(fmaxf(0.f, dy0), dy1);
  const float A00 = a1x * a1x + a2x * a2x, A01 = a1x * a1y + a2x * a2y, A11 = a1y * a1y + a2y * a2y;
  const float dyv = fminf(fmaxf(-A01 * dxc * fast_rcp(A11), dy0), dy1);
  const float qv = A00 * dxc * dxc + 2.f * A01 * dxc * dyv + A11 * dyv * dyv;
  const float dxh = fminf(fmaxf(-A01 * dyc * fast_rcp(A00), dx0), dx1);
  const float qh = A00 * dxh * dxh + 2.f * A01 * dxh * dyc + A11 * dyc * dyc;
  return fminf(qv, qh) < qmax * 1.001f + 1e-3f;
}

// ------------------------------------------------------------------------------------------------ warp reductions
// Transposed butterfly over NV per-lane partial sums: every exchange step halves the number of live values
// (the lane keeps the half selected by its lane bit and adds the partner's copy of it), so NV values cost about
// NV shuffles instead of 5 NV.  reduce_owner<NV>(lane) tells which value ends up, fully summed, in v[0] of a lane.
template <int N, int OFF>
__device__ __forceinline__ void reduce_scatter_step(float* v, int lane) {
  if constexpr (OFF >= 1) {
    if constexpr (N > 1) {
      constexpr int H = (N + 1) / 2;
      const bool upper = (lane & OFF) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        if (i + H < N) {
          const float send = upper ? v[i] : v[i + H];
          const float keep = upper ? v[i + H] : v[i];
          v[i] = keep + __shfl_xor_sync(kFull, send, OFF);
        } else {
          // odd N: the unpaired value is summed on both sides (no selects); the upper half's copy is a duplicate
          // that reduce_owner never hands out
          v[i] += __shfl_xor_sync(kFull, v[i], OFF);
        }
      }
      reduce_scatter_step<H, OFF / 2>(v, lane);
    } else {
      v[0] += __shfl_xor_sync(kFull, v[0], OFF);
      reduce_scatter_step<1, OFF / 2>(v, lane);
    }
  }
}

template <int NV>
__device__ __forceinline__ int reduce_owner(int lane) {
  int base = 0, cnt = NV, n = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int half = (n + 1) / 2;
      if (lane & off) { base += half; cnt = max(cnt - half, 0); }
      else cnt = min(cnt, half);
      n = half;
    } else if (lane & off) {
      cnt = 0;
    }
  }
  return cnt == 1 ? base : -1;
}

int raster_fast_pack(const GsRasterParams& p, const RasterArgs& a, bool forward, bool features, cudaStream_t st);
// cull masks of every tile-list entry against the eight 8x4 pixel blocks of its tile (needs the forward records)
int raster_cull_mask(const GsRasterParams& p, const RasterArgs& a, cudaStream_t st);

// raster_fast_bwd_wide.cu: backward for 8..64 feature channels (one pixel per lane, image gradient in registers)
int raster_bwd_wide(const GsRasterParams& p, const RasterArgs& a, const float4* rec, const float* featP, cudaStream_t st);

}  // namespace gs
