// raster_fast_bwd.cu — backward rasterizer, f32 / alpha blending / tile 16 (the measured path).
//
// Replaces rasterizer/backward.py:52-228 of /root/reference/taichi_splatting/.  Same per pixel arithmetic:
// front-to-back replay that subtracts each blended term from the final forward image (R -= f w),
// dL/dalpha = sum_c (f_c T - R_c / (1 - alpha)) G_c, threshold tested on the unclamped alpha and gradient
// passed through the clamp, pixels stop at saturate_threshold.  Machine mapping:
//   * one CTA per 16x16 tile, 2 warps, each warp owns a 16x8 pixel block, each lane a 2x2 quad (the
//     reference's default pixel_stride; the quad amortises the per gaussian warp reduction over 4 pixels);
//   * cp.async double-buffered staging of 32 B records + padded feature rows, batches of 64;
//   * lane-parallel ellipse / pixel-block cull before the per pixel work (see raster_fast.cuh);
//   * the 7 + F (+2) per gaussian partial sums are reduced across the warp with a transposed butterfly
//     (16 shuffles for up to 16 values instead of 5 per value); the lane that ends up owning a value
//     issues one red.global.add for it — one reduction per (gaussian, warp), no shared-memory atomics
//     and no block barrier inside the batch.
// Per pixel state lives in registers; F is a template parameter (1..7 here, wider F takes raster_generic.cu).
#include "raster_fast.cuh"

namespace gs {

constexpr int kBwdBatch = 64;
constexpr int kBwdThreads = 64;

// Transposed butterfly: on return v[0] of lane L holds the warp total of input value (L >> 1).
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane) {
  static_assert(NV == 16, "16 values");
#pragma unroll
  for (int half = NV / 2, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(kFull, send, off);
    }
  }
  v[0] += __shfl_xor_sync(kFull, v[0], 1);
}

template <int F, int FP, bool HEUR>
__global__ void __launch_bounds__(kBwdThreads)
raster_bwd_fast_kernel(const __grid_constant__ GsRasterParams p, const float4* __restrict__ rec,
                       const float* __restrict__ featP, const int32_t* __restrict__ ranges,
                       const int32_t* __restrict__ o2p, const float* __restrict__ image,
                       const float* __restrict__ grad_image, float* __restrict__ grad_pts,
                       float* __restrict__ grad_feat, float* __restrict__ heuristic) {
  __shared__ __align__(16) float4 s_r0[2][kBwdBatch];
  __shared__ __align__(16) float4 s_r1[2][kBwdBatch];
  __shared__ __align__(16) float s_feat[2][kBwdBatch][FP];

  const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tw = (p.image_width + kFastTile - 1) / kFastTile;
  const int ox = (tile % tw) * kFastTile, oy = (tile / tw) * kFastTile + warp * 8;
  const int x0 = ox + 2 * (lane & 7), y0 = oy + 2 * (lane >> 3);
  const float bx0 = (float)ox + 0.5f, bx1 = (float)ox + 15.5f, by0 = (float)oy + 0.5f, by1 = (float)oy + 7.5f;
  const float thr = (float)p.alpha_threshold, cmax = (float)p.clamp_max_alpha, sat = (float)p.saturate_threshold;
  const float l2thr = log2f(thr);
  const bool pg = p.points_requires_grad && grad_pts != nullptr;
  const bool fg = p.features_requires_grad && grad_feat != nullptr;

  float W[4], R[4][F], Gd[4][F], pxf[4], pyf[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int px = x0 + (i & 1), py = y0 + (i >> 1);
    pxf[i] = (float)px + 0.5f; pyf[i] = (float)py + 0.5f;
    const bool inb = px < p.image_width && py < p.image_height;
    W[i] = inb ? 0.f : 1.f;
    const int64_t pix = (int64_t)py * p.image_width + px;
#pragma unroll
    for (int c = 0; c < F; ++c) {
      R[i][c] = inb ? image[pix * F + c] : 0.f;
      Gd[i][c] = inb ? grad_image[pix * F + c] : 0.f;
    }
  }

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  const int C = end - start;
  const int nb = (C + kBwdBatch - 1) / kBwdBatch;

  auto issue_load = [&](int b) {
    const int buf = b & 1;
    const int v = b * kBwdBatch + t;
    if (v < C) {
      const int idx = o2p[start + v];
      cp_async16(&s_r0[buf][t], rec + 2 * (int64_t)idx);
      cp_async16(&s_r1[buf][t], rec + 2 * (int64_t)idx + 1);
#pragma unroll
      for (int c = 0; c < FP; c += 4) cp_async16(&s_feat[buf][t][c], featP + (int64_t)idx * FP + c);
    }
    cp_async_commit();
  };

  auto lane_done = [&]() { return W[0] >= sat && W[1] >= sat && W[2] >= sat && W[3] >= sat; };
  bool warp_done = __all_sync(kFull, lane_done());
  if (nb > 0) issue_load(0);
  for (int b = 0; b < nb; ++b) {
    const int buf = b & 1;
    if (b + 1 < nb) {
      issue_load(b + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int n_in = min(kBwdBatch, C - b * kBwdBatch);
    if (!warp_done) {
      for (int c0 = 0; c0 < n_in; c0 += 32) {
        const int e = c0 + lane;
        bool hit = false;
        if (e < n_in) {
          const float4 r0 = s_r0[buf][e], r1 = s_r1[buf][e];
          const float cx = kSqrtHalfLog2e * r1.x, cy = kSqrtHalfLog2e * r1.y;
          hit = block_may_touch(r0.x, r0.y, r0.z * cx, r0.w * cx, -r0.w * cy, r0.z * cy, log2f(r1.z) - l2thr, bx0,
                                bx1, by0, by1);
        }
        unsigned mask = __ballot_sync(kFull, hit);
        while (mask) {
          const int j = c0 + __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 r0 = s_r0[buf][j], r1 = s_r1[buf][j];
          const float mx = r0.x, my = r0.y, ax = r0.z, ay = r0.w, isx = r1.x, isy = r1.y, a0 = r1.z;
          float f[F];
#pragma unroll
          for (int c = 0; c < F; ++c) f[c] = s_feat[buf][j][c];

          float U = 0.f, V = 0.f, Sx = 0.f, Sy = 0.f, Ax = 0.f, Ay = 0.f, Ga = 0.f, h0 = 0.f, h1 = 0.f;
          float gf[F];
#pragma unroll
          for (int c = 0; c < F; ++c) gf[c] = 0.f;
          bool has_grad = false;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float dx = pxf[i] - mx, dy = pyf[i] - my;
            const float tx = fmaf(dy, ay, dx * ax) * isx;
            const float ty = fmaf(dy, ax, -dx * ay) * isy;
            const float q = fmaf(ty, ty, tx * tx);
            const float pgauss = fast_ex2(-kHalfLog2e * q);
            float alpha = a0 * pgauss;
            if (alpha > thr && W[i] < sat) {
              has_grad = true;
              alpha = fminf(alpha, cmax);
              const float Ti = 1.f - W[i];
              const float w = alpha * Ti;
              W[i] += w;
              const float rinv = fast_rcp(1.f - alpha);
              float ag = 0.f;
#pragma unroll
              for (int c = 0; c < F; ++c) {
                R[i][c] = fmaf(-f[c], w, R[i][c]);
                const float diff = fmaf(f[c], Ti, -R[i][c] * rinv);
                ag = fmaf(diff, Gd[i][c], ag);
                gf[c] = fmaf(w, Gd[i][c], gf[c]);
              }
              const float aag = a0 * ag;
              const float g = aag * pgauss;
              const float a = g * tx * isx, bq = g * ty * isy;
              U += a; V += bq;
              Sx = fmaf(a, tx, Sx); Sy = fmaf(bq, ty, Sy);
              Ax -= fmaf(a, dx, bq * dy);
              Ay += fmaf(bq, dx, -a * dy);
              Ga = fmaf(pgauss, ag, Ga);
              if (HEUR) {
                h0 = fmaf(aag, aag, h0);
                h1 += fabsf(fmaf(a, ax, -bq * ay)) + fabsf(fmaf(a, ay, bq * ax));
              }
            }
          }
          if (__any_sync(kFull, has_grad)) {
            float v[16];
            v[0] = fmaf(ax, U, -ay * V); v[1] = fmaf(ay, U, ax * V);
            v[2] = Ax; v[3] = Ay; v[4] = Sx; v[5] = Sy; v[6] = Ga;
#pragma unroll
            for (int c = 0; c < F; ++c) v[7 + c] = gf[c];
            if (HEUR) { v[7 + F] = h0; v[8 + F] = h1; }
#pragma unroll
            for (int k = 7 + F + (HEUR ? 2 : 0); k < 16; ++k) v[k] = 0.f;
            warp_reduce_scatter<16>(v, lane);
            const int vi = lane >> 1;
            const int64_t idx = __float_as_int(r1.w);
            if ((lane & 1) == 0) {
              if (vi < 7) { if (pg) atomicAdd(grad_pts + idx * 7 + vi, v[0]); }
              else if (vi < 7 + F) { if (fg) atomicAdd(grad_feat + idx * F + (vi - 7), v[0]); }
              else if (HEUR && vi < 9 + F) atomicAdd(heuristic + idx * 2 + (vi - 7 - F), v[0]);
            }
          }
        }
      }
      warp_done = __all_sync(kFull, lane_done());
    }
    if (__syncthreads_and(warp_done)) break;
  }
  cp_async_wait<0>();
}

template <int F, int FP>
static int launch_bwd_fast(const GsRasterParams& p, const RasterArgs& a, const float4* rec, const float* featP,
                           cudaStream_t st) {
  const int tiles = tiles_wide(p) * tiles_high(p);
  const bool heur = p.compute_point_heuristic && a.point_heuristic != nullptr;
  if (heur)
    raster_bwd_fast_kernel<F, FP, true><<<tiles, kBwdThreads, 0, st>>>(
        p, rec, featP, a.tile_ranges, a.overlap_to_point, (const float*)a.image_in, (const float*)a.grad_image,
        (float*)a.grad_gaussians, (float*)a.grad_features, (float*)a.point_heuristic);
  else
    raster_bwd_fast_kernel<F, FP, false><<<tiles, kBwdThreads, 0, st>>>(
        p, rec, featP, a.tile_ranges, a.overlap_to_point, (const float*)a.image_in, (const float*)a.grad_image,
        (float*)a.grad_gaussians, (float*)a.grad_features, nullptr);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

bool raster_bwd_fast_supported(const GsRasterParams& p) { return raster_fast_supported(p) && p.num_features <= 7; }

int raster_bwd_fast(const GsRasterParams& p, const RasterArgs& a, cudaStream_t st) {
  if (p.num_features > 7) return raster_bwd_generic(p, a, st);
  int rc = raster_fast_pack(p, a, /*forward=*/false, /*features=*/!p.workspace_holds_packed, st);
  if (rc != GS_OK) return rc;
  const FastLayout L = fast_layout(p);
  unsigned char* ws = (unsigned char*)a.workspace;
  const float4* rec = (const float4*)(ws + L.off_recB);
  const float* featP = (const float*)(ws + L.off_feat);
  switch (p.num_features) {
    case 1: return launch_bwd_fast<1, 4>(p, a, rec, featP, st);
    case 2: return launch_bwd_fast<2, 4>(p, a, rec, featP, st);
    case 3: return launch_bwd_fast<3, 4>(p, a, rec, featP, st);
    case 4: return launch_bwd_fast<4, 4>(p, a, rec, featP, st);
    case 5: return launch_bwd_fast<5, 8>(p, a, rec, featP, st);
    case 6: return launch_bwd_fast<6, 8>(p, a, rec, featP, st);
    case 7: return launch_bwd_fast<7, 8>(p, a, rec, featP, st);
    default: break;
  }
  GS_UNSUPPORTED("rasterizer backward (fast): %d feature channels", p.num_features);
}

}  // namespace gs
