// geom_kernels.cu — projection forward (fused cull + ordered compaction), tile overlap count,
// sort-key emission, tile range detection.   Build with -fmad=false (bit-exactness contract,
// see include/gs_numeric.h).
//
// Replaces (paths relative to /root/reference/taichi_splatting/):
//   project_kernel + nonzero + gathers      perspective/projection.py:31-80, :146-149
//   tile_overlaps_kernel                    mapper/tile_mapper.py:73-84
//   generate_sort_keys_kernel               mapper/tile_mapper.py:112-144
//   find_ranges_kernel                      mapper/tile_mapper.py:90-110
#include "common.cuh"
#include "geom_math.cuh"
#include "lookback.cuh"

namespace gs {

// ------------------------------------------------------------------------------------------------
// Projection forward.  One thread per gaussian; culled gaussians are dropped in-kernel with an
// order-preserving compaction (block scan + decoupled look-back), so `indexes` comes out ascending
// exactly like torch.nonzero() without materialising the (N,7) intermediate or syncing the host.
// HBM traffic: 44 B read per gaussian, 44 B written per visible gaussian.
// ------------------------------------------------------------------------------------------------
constexpr int kProjBlock = 256;

template <typename T>
__device__ __forceinline__ void load_camera(const T* __restrict__ Tcw, const T* __restrict__ proj,
                                            const GsProjectParams& p, CameraConst<T>& C) {
#pragma unroll
  for (int i = 0; i < 12; ++i) C.Tcw[i] = Tcw[i];
  C.fx = proj[0]; C.fy = proj[1]; C.cx = proj[2]; C.cy = proj[3];
  C.w = T(p.image_width); C.h = T(p.image_height);
  C.near_ = T(p.near_plane); C.far_ = T(p.far_plane); C.blur = T(p.blur_cov);
  C.lo_x = T(-(double)p.image_width * p.clamp_margin);
  C.lo_y = T(-(double)p.image_height * p.clamp_margin);
  C.hi_x = T(((double)p.image_width - 1.0) * (1.0 + p.clamp_margin));
  C.hi_y = T(((double)p.image_height - 1.0) * (1.0 + p.clamp_margin));
  C.alpha_threshold = T(p.alpha_threshold);
}

// kProjItems gaussians per thread (slot k of thread t is gaussian tile * kProjTile + k * kProjBlock + t, so loads stay
// coalesced): a quarter of the look-back links of a one-gaussian-per-thread layout, and four independent projections
// in flight per thread while the chain resolves (the 11.7 K link chain was the critical path: 0.19 ms at 39 % issue).
constexpr int kProjItems = 4;
constexpr int kProjTile = kProjBlock * kProjItems;
static_assert(kProjItems * (kProjBlock / 32) == 32, "one warp scans the (item, warp) counts");

template <typename T>
__global__ void __launch_bounds__(kProjBlock, sizeof(T) == 4 ? 6 : 1)   // 4 / 5 / 6 / 8 CTAs per SM measured: 0.141 / 0.135 / 0.131 / 0.142 ms
project_fwd_kernel(const __grid_constant__ GsProjectParams p, const T* __restrict__ position,
                   const T* __restrict__ log_scaling, const T* __restrict__ rotation,
                   const T* __restrict__ alpha_logit, const T* __restrict__ Tcw, const T* __restrict__ proj,
                   T* __restrict__ points, T* __restrict__ depth, int64_t* __restrict__ indexes,
                   int32_t* __restrict__ num_visible, unsigned long long* __restrict__ status,
                   unsigned int* __restrict__ ticket) {
  __shared__ int s_tile;
  __shared__ int s_count[32];  // [item][warp]: visible gaussians, then their exclusive offsets in the tile
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * kProjTile + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  CameraConst<T> C;
  load_camera<T>(Tcw, proj, p, C);

  Projected<T> o[kProjItems];
  unsigned ballot[kProjItems];
#pragma unroll
  for (int k = 0; k < kProjItems; ++k) {
    const int64_t i = base + (int64_t)k * kProjBlock;
    o[k].in_view = false;
    if (i < p.num_points) {
      T pos[3] = {position[3 * i], position[3 * i + 1], position[3 * i + 2]};
      T ls[3] = {log_scaling[3 * i], log_scaling[3 * i + 1], log_scaling[3 * i + 2]};
      T q[4];
      if (sizeof(T) == 4) {
        float4 qv = reinterpret_cast<const float4*>(rotation)[i];
        q[0] = qv.x; q[1] = qv.y; q[2] = qv.z; q[3] = qv.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = rotation[4 * i + c];
      }
      o[k] = project_one<T>(pos, ls, q, alpha_logit[i], C);
    }
    ballot[k] = __ballot_sync(kFull, o[k].in_view);
    if (lane == 0) s_count[k * (kProjBlock / 32) + warp] = __popc(ballot[k]);
  }
  __syncthreads();
  if (warp == 0) {
    const int c = s_count[lane];
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    s_count[lane] = incl - c;
    const int block_total = __shfl_sync(kFull, incl, 31);
    unsigned long long ex = lookback_exclusive(status, tile, (unsigned long long)block_total);
    if (lane == 0) {
      s_prefix = ex;
      if (tile == (int)gridDim.x - 1) *num_visible = (int32_t)(ex + block_total);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kProjItems; ++k) {
    if (o[k].in_view) {
      const int64_t dst = (int64_t)s_prefix + s_count[k * (kProjBlock / 32) + warp] +
                          __popc(ballot[k] & ((1u << lane) - 1u));
      T* g = points + 7 * dst;
      g[0] = o[k].mean_x; g[1] = o[k].mean_y; g[2] = o[k].axis_x; g[3] = o[k].axis_y;
      g[4] = o[k].sigma_x; g[5] = o[k].sigma_y; g[6] = o[k].alpha;
      depth[dst] = o[k].z;
      indexes[dst] = base + (int64_t)k * kProjBlock;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile overlap count / key emission.  One lane per gaussian computes the OBB query; spans of at
// most kSerialSpan tiles are walked by the owning lane, larger spans are handed to the whole warp
// (query broadcast by shuffle, lanes stride over the span, ballot/popc for counts and ranks) so a
// single screen-filling gaussian does not stall 31 idle lanes.  The predicate is the same function
// in both passes.  Emission order within a gaussian is x-outer / y-inner as in the reference.
// ------------------------------------------------------------------------------------------------
constexpr int kTileBlock = 128;
constexpr int kSerialSpan = 32;

__device__ __forceinline__ uint64_t make_key64(float depth, int tile_id) {
  return (uint64_t)__float_as_uint(depth) | ((uint64_t)(uint32_t)tile_id << 32);
}
__device__ __forceinline__ uint32_t make_key32(float depth, int tile_id) {
  float c = depth < 0.f ? 0.f : (depth > 1.f ? 1.f : depth);
  return (uint32_t)(c * 65535.0f) | ((uint32_t)tile_id << 16);
}

__device__ __forceinline__ TileQuery shfl_query(const TileQuery& q, int src) {
  TileQuery r;
  r.ib00 = __shfl_sync(kFull, q.ib00, src); r.ib01 = __shfl_sync(kFull, q.ib01, src);
  r.ib10 = __shfl_sync(kFull, q.ib10, src); r.ib11 = __shfl_sync(kFull, q.ib11, src);
  r.rel_x = __shfl_sync(kFull, q.rel_x, src); r.rel_y = __shfl_sync(kFull, q.rel_y, src);
  r.min_x = __shfl_sync(kFull, q.min_x, src); r.min_y = __shfl_sync(kFull, q.min_y, src);
  r.span_x = __shfl_sync(kFull, q.span_x, src); r.span_y = __shfl_sync(kFull, q.span_y, src);
  return r;
}

// MODE 0: counts[i] = number of overlapped tiles.  MODE 1: write (tile, depth) keys / values at cum[i].
// MODE 2: write bare tile ids (u32) / values at cum[i] (depth-first pipeline: the gaussians are visited in depth
// order through `perm`, so a stable sort on the tile id alone yields the (tile, depth, index) order).
// perm (optional): slot i works on gaussian perm[i]; values always hold the gaussian index.
// masks (optional, one uint2 per slot): the count pass records which tiles of a span of at most 32 passed the test
// (.x: bit tu * span_y + tv; .y: min_x | min_y << 12 | span_y << 24, bit 31 = span too large, recompute), and the emit
// pass expands the bits instead of evaluating the OBB query and up to 32 tile tests a second time.
constexpr unsigned kMaskRecompute = 1u << 31;

template <int MODE, typename KeyT>
__global__ void __launch_bounds__(kTileBlock)
tile_query_kernel(const __grid_constant__ GsTileParams p, int img_w, int img_h, const float* __restrict__ g,
                  const float* __restrict__ depth, const int32_t* __restrict__ perm, const int32_t* __restrict__ cum,
                  int32_t* __restrict__ counts, KeyT* __restrict__ keys, int32_t* __restrict__ values,
                  uint2* __restrict__ masks, const int32_t* __restrict__ n_dev = nullptr,
                  long long limit = 0x7fffffffffffffffll) {
  // limit: capacity of keys / values (emit passes): entries that would land at or beyond it are dropped, so a
  // capacity-sized buffer (no host read-back of the overlap total, CUDA-graph friendly) cannot be overrun
  constexpr bool EMIT = MODE != 0;
  const int64_t slot = (int64_t)blockIdx.x * kTileBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int ts = p.tile_size;
  const int tiles_wide = img_w / ts;
  // counted variant (n_dev): p.num_points is the capacity, the number of gaussians is still on the device
  const bool valid = slot < (n_dev ? min(p.num_points, (int64_t)*n_dev) : p.num_points);

  TileQuery qy;
  qy.span_x = qy.span_y = 0;
  float d = 0.f;
  int64_t base = 0;
  int64_t i = slot;
  bool from_mask = false;
  if (valid) {
    if (perm) i = perm[slot];
    if (EMIT) base = cum[slot];
    if (MODE == 2 && masks != nullptr) {
      const uint2 m = masks[slot];
      if (!(m.y & kMaskRecompute)) {  // expand the recorded bits, in bit order = x outer / y inner
        from_mask = true;
        const int min_x = m.y & 0xfff, min_y = (m.y >> 12) & 0xfff, span_y = (m.y >> 24) & 0x3f;
        unsigned bits = m.x;
        int c = 0;
        // b / span_y for b < 32, span_y <= 32 without the integer division sequence: (b + 0.5) / span_y stays at
        // least 1 / 64 away from every integer, far more than the rounding of the float product
        const float inv_span = 1.0f / (float)span_y;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int tu = (int)(((float)b + 0.5f) * inv_span), tv = b - tu * span_y;
          if (base + c < limit) {
            keys[base + c] = (KeyT)((tu + min_x) + (tv + min_y) * tiles_wide);
            values[base + c] = (int32_t)i;
          }
          ++c;
        }
      }
    }
    if (!from_mask) {
      const float* gi = g + 7 * i;
      qy = obb_query(gi[0], gi[1], gi[2], gi[3], gi[4], gi[5], gi[6], img_w, img_h, ts, (float)p.alpha_threshold);
      if (MODE == 1) d = depth[i];
    }
  }
  const int n = span_count(qy);
  int count = 0;

  if (n <= kSerialSpan) {
    unsigned bits = 0;
    for (int tu = 0; tu < qy.span_x; ++tu)
      for (int tv = 0; tv < qy.span_y; ++tv)
        if (test_tile(qy, tu, tv, ts)) {
          if (EMIT && base + count < limit) {
            int tile_id = (tu + qy.min_x) + (tv + qy.min_y) * tiles_wide;
            keys[base + count] = MODE == 2 ? (KeyT)tile_id
                                 : sizeof(KeyT) == 8 ? (KeyT)make_key64(d, tile_id) : (KeyT)make_key32(d, tile_id);
            values[base + count] = (int32_t)i;
          } else {
            bits |= 1u << (tu * qy.span_y + tv);
          }
          ++count;
        }
    if (!EMIT && masks != nullptr && valid) {
      const bool fits = qy.min_x < 4096 && qy.min_y < 4096;
      masks[slot] = fits ? make_uint2(bits, (unsigned)qy.min_x | ((unsigned)qy.min_y << 12) | ((unsigned)qy.span_y << 24))
                         : make_uint2(0u, kMaskRecompute);
    }
  } else if (!EMIT && masks != nullptr && valid) {
    masks[slot] = make_uint2(0u, kMaskRecompute);
  }
  unsigned big = __ballot_sync(kFull, n > kSerialSpan);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const TileQuery q = shfl_query(qy, src);
    const int nq = q.span_x * q.span_y;
    const float dq = __shfl_sync(kFull, d, src);
    const long long bq = __shfl_sync(kFull, (long long)base, src);
    const long long iq = __shfl_sync(kFull, (long long)i, src);
    int running = 0;
    for (int t0 = 0; t0 < nq; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      int tu = 0, tv = 0;
      if (t < nq) {
        tu = t / q.span_y; tv = t - tu * q.span_y;
        hit = test_tile(q, tu, tv, ts);
      }
      const unsigned hm = __ballot_sync(kFull, hit);
      if (EMIT && hit && bq + running + __popc(hm & ((1u << lane) - 1u)) < limit) {
        const int r = running + __popc(hm & ((1u << lane) - 1u));
        int tile_id = (tu + q.min_x) + (tv + q.min_y) * tiles_wide;
        keys[bq + r] = MODE == 2 ? (KeyT)tile_id
                       : sizeof(KeyT) == 8 ? (KeyT)make_key64(dq, tile_id) : (KeyT)make_key32(dq, tile_id);
        values[bq + r] = (int32_t)iq;
      }
      running += __popc(hm);
    }
    if (lane == src) count = running;
  }
  if (!EMIT && valid) counts[slot] = count;
}

// ------------------------------------------------------------------------------------------------
// Tile ranges from sorted keys.  ranges[t] = [first, last+1); tiles without overlaps stay [0,0].
// ------------------------------------------------------------------------------------------------
template <typename KeyT, int SHIFT>
__global__ void find_ranges_kernel(int64_t n, const KeyT* __restrict__ keys, int32_t* __restrict__ ranges,
                                   const int32_t* __restrict__ n_dev = nullptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, (int64_t)*n_dev);   // counted variant: n is the capacity
  if (i >= n) return;
  const int max_tile = 65535;
  const int tile = (int)(keys[i] >> SHIFT);
  int next = max_tile;
  if (i + 1 < n) next = (int)(keys[i + 1] >> SHIFT);
  if (tile != next) {
    ranges[2 * tile + 1] = (int32_t)(i + 1);
    if (next < max_tile) ranges[2 * next] = (int32_t)(i + 1);
  }
}

// u32 sort keys of the depth-first pipeline: the f32 bit pattern of the depth (mapper/tile_mapper.py:34-40) or
// the 16 bit quantisation (:53-59); values = gaussian index.
// to_ndc: the depth is linear camera depth and the sort depth is its NDC value, computed with exactly the four f32
// operations torch's eager CUDA kernels perform for  1 - (1/d - 1/far) / (1/near - 1/far)  (reciprocal, subtract,
// multiply by the reciprocal of the scalar divisor, subtract from one), so the keys carry the same bits as
// ndc_depth() on the device.  This translation unit is compiled with -fmad=false.
__global__ void depth_keys_kernel(int64_t n, int use_depth16, int to_ndc, float inv_far, float inv_den,
                                  const float* __restrict__ depth, uint32_t* __restrict__ keys,
                                  int32_t* __restrict__ values, const int32_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, (int64_t)*n_dev);   // counted variant: n is the capacity
  if (i >= n) return;
  float d = depth[i];
  if (to_ndc) d = 1.0f - ((1.0f / d) - inv_far) * inv_den;
  keys[i] = use_depth16 ? (make_key32(d, 0) & 0xffffu) : (uint32_t)(make_key64(d, 0) & 0xffffffffull);
  values[i] = (int32_t)i;
}

// ---- camera centre: -A^-1 t of the affine view matrix [A | t], A^-1 from cross products (one thread; replaces the
// ten ATen launches of the same closed form, or torch.inverse + its host sync, perspective/params.py:75-78)
template <typename T>
__global__ void camera_position_kernel(const T* __restrict__ M, T* __restrict__ out) {
  const T a0x = M[0], a0y = M[1], a0z = M[2], tx = M[3];
  const T a1x = M[4], a1y = M[5], a1z = M[6], ty = M[7];
  const T a2x = M[8], a2y = M[9], a2z = M[10], tz = M[11];
  const T c0x = a1y * a2z - a1z * a2y, c0y = a1z * a2x - a1x * a2z, c0z = a1x * a2y - a1y * a2x;   // a1 x a2
  const T c1x = a2y * a0z - a2z * a0y, c1y = a2z * a0x - a2x * a0z, c1z = a2x * a0y - a2y * a0x;   // a2 x a0
  const T c2x = a0y * a1z - a0z * a1y, c2y = a0z * a1x - a0x * a1z, c2z = a0x * a1y - a0y * a1x;   // a0 x a1
  const T det = a0x * c0x + a0y * c0y + a0z * c0z;
  out[0] = -(c0x * tx + c1x * ty + c2x * tz) / det;
  out[1] = -(c0y * tx + c1y * ty + c2y * tz) / det;
  out[2] = -(c0z * tx + c1z * ty + c2z * tz) / det;
}

}  // namespace gs

using namespace gs;

// ================================================================================================ C ABI
extern "C" {

size_t gs_project_fwd_workspace_bytes(const GsProjectParams* p) {
  if (!p) return 0;
  int64_t blocks = ceil_div(p->num_points > 0 ? p->num_points : 1, kProjTile);
  return align_up((size_t)blocks * sizeof(unsigned long long) + 16, 256);
}

int gs_project_fwd(const GsProjectParams* p, const void* position, const void* log_scaling, const void* rotation,
                   const void* alpha_logit, const void* T_camera_world, const void* projection, void* points,
                   void* depth, int64_t* indexes, int32_t* num_visible, void* workspace, size_t workspace_bytes,
                   void* stream) {
  GS_CHECK_ARG(p != nullptr, "gs_project_fwd: null params");
  GS_CHECK_ARG(p->dtype == GS_F32 || p->dtype == GS_F64, "gs_project_fwd: dtype must be GS_F32 or GS_F64");
  GS_CHECK_ARG(p->num_points >= 0 && p->image_width > 0 && p->image_height > 0, "gs_project_fwd: bad sizes");
  GS_CHECK_ARG(num_visible != nullptr, "gs_project_fwd: num_visible is null");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->num_points == 0) {
    GS_CUDA(cudaMemsetAsync(num_visible, 0, sizeof(int32_t), st));
    return GS_OK;
  }
  GS_CHECK_ARG(position && log_scaling && rotation && alpha_logit && T_camera_world && projection && points &&
                   depth && indexes, "gs_project_fwd: null tensor");
  size_t need = gs_project_fwd_workspace_bytes(p);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gs_project_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return GS_ERR_WORKSPACE;
  }
  int64_t blocks = ceil_div(p->num_points, kProjTile);
  GS_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  unsigned long long* status = (unsigned long long*)workspace;
  unsigned int* ticket = (unsigned int*)(status + blocks);
  if (p->dtype == GS_F32) {
    project_fwd_kernel<float><<<(unsigned)blocks, kProjBlock, 0, st>>>(
        *p, (const float*)position, (const float*)log_scaling, (const float*)rotation, (const float*)alpha_logit,
        (const float*)T_camera_world, (const float*)projection, (float*)points, (float*)depth, indexes, num_visible,
        status, ticket);
  } else {
    project_fwd_kernel<double><<<(unsigned)blocks, kProjBlock, 0, st>>>(
        *p, (const double*)position, (const double*)log_scaling, (const double*)rotation,
        (const double*)alpha_logit, (const double*)T_camera_world, (const double*)projection, (double*)points,
        (double*)depth, indexes, num_visible, status, ticket);
  }
  GS_LAUNCH_CHECK();
  return GS_OK;
}

static int check_tile_params(const GsTileParams* p, const char* who) {
  if (!p) { set_error("%s: null params", who); return GS_ERR_INVALID; }
  if (p->tile_size <= 0 || p->image_width <= 0 || p->image_height <= 0 || p->num_points < 0) {
    set_error("%s: bad sizes", who);
    return GS_ERR_INVALID;
  }
  int64_t tw = ceil_div(p->image_width, p->tile_size), th = ceil_div(p->image_height, p->tile_size);
  if (tw * th >= 65535) {  // mapper/tile_mapper.py:175-176
    set_error("%s: %lld x %lld tiles exceed the 16 bit tile id, increase tile_size", who, (long long)th,
              (long long)tw);
    return GS_ERR_INVALID;
  }
  return GS_OK;
}

int gs_tile_count(const GsTileParams* p, const float* gaussians, int32_t* counts, void* stream) {
  int rc = check_tile_params(p, "gs_tile_count");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0) return GS_OK;
  GS_CHECK_ARG(gaussians && counts, "gs_tile_count: null tensor");
  int ts = p->tile_size;
  int img_w = (int)ceil_div(p->image_width, ts) * ts, img_h = (int)ceil_div(p->image_height, ts) * ts;
  int64_t blocks = ceil_div(p->num_points, kTileBlock);
  tile_query_kernel<0, uint64_t><<<(unsigned)blocks, kTileBlock, 0, (cudaStream_t)stream>>>(
      *p, img_w, img_h, gaussians, nullptr, nullptr, nullptr, counts, nullptr, nullptr, nullptr);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_tile_emit_keys(const GsTileParams* p, const float* gaussians, const float* depth, const int32_t* cum,
                      void* keys, int32_t* values, void* stream) {
  int rc = check_tile_params(p, "gs_tile_emit_keys");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0) return GS_OK;
  GS_CHECK_ARG(gaussians && depth && cum && keys && values, "gs_tile_emit_keys: null tensor");
  int ts = p->tile_size;
  int img_w = (int)ceil_div(p->image_width, ts) * ts, img_h = (int)ceil_div(p->image_height, ts) * ts;
  int64_t blocks = ceil_div(p->num_points, kTileBlock);
  if (p->use_depth16)
    tile_query_kernel<1, uint32_t><<<(unsigned)blocks, kTileBlock, 0, (cudaStream_t)stream>>>(
        *p, img_w, img_h, gaussians, depth, nullptr, cum, nullptr, (uint32_t*)keys, values, nullptr);
  else
    tile_query_kernel<1, uint64_t><<<(unsigned)blocks, kTileBlock, 0, (cudaStream_t)stream>>>(
        *p, img_w, img_h, gaussians, depth, nullptr, cum, nullptr, (uint64_t*)keys, values, nullptr);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_find_ranges(const GsTileParams* p, int64_t num_overlaps, const void* sorted_keys, int32_t* tile_ranges,
                   void* stream) {
  int rc = check_tile_params(p, "gs_find_ranges");
  if (rc != GS_OK) return rc;
  GS_CHECK_ARG(tile_ranges != nullptr && num_overlaps >= 0, "gs_find_ranges: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t tiles = ceil_div(p->image_width, p->tile_size) * ceil_div(p->image_height, p->tile_size);
  GS_CUDA(cudaMemsetAsync(tile_ranges, 0, (size_t)tiles * 2 * sizeof(int32_t), st));
  if (num_overlaps == 0) return GS_OK;
  GS_CHECK_ARG(sorted_keys != nullptr, "gs_find_ranges: null keys");
  int64_t blocks = ceil_div(num_overlaps, 256);
  if (p->use_depth16)
    find_ranges_kernel<uint32_t, 16><<<(unsigned)blocks, 256, 0, st>>>(num_overlaps, (const uint32_t*)sorted_keys,
                                                                      tile_ranges);
  else
    find_ranges_kernel<uint64_t, 32><<<(unsigned)blocks, 256, 0, st>>>(num_overlaps, (const uint64_t*)sorted_keys,
                                                                      tile_ranges);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_camera_position(int32_t dtype, const void* T_camera_world, void* position, void* stream) {
  GS_CHECK_ARG(T_camera_world && position, "gs_camera_position: null tensor");
  GS_CHECK_ARG(dtype == GS_F32 || dtype == GS_F64, "gs_camera_position: dtype must be GS_F32 or GS_F64");
  if (dtype == GS_F32)
    camera_position_kernel<float><<<1, 1, 0, (cudaStream_t)stream>>>((const float*)T_camera_world, (float*)position);
  else
    camera_position_kernel<double><<<1, 1, 0, (cudaStream_t)stream>>>((const double*)T_camera_world, (double*)position);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- depth-first tile mapping (same outputs as count -> emit_keys -> sort on 32 + tile bits -> find_ranges):
// sort the V gaussians by depth key once, visit them in that order, then a stable sort on the tile id only.
static int depth_keys_entry(const GsTileParams* p, const float* depth, double near_plane, double far_plane,
                            const int32_t* count_dev, uint32_t* keys, int32_t* values, void* stream) {
  int rc = check_tile_params(p, "gs_depth_keys");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0) return GS_OK;
  GS_CHECK_ARG(depth && keys && values, "gs_depth_keys: null tensor");
  const bool to_ndc = near_plane > 0.0;
  GS_CHECK_ARG(!to_ndc || far_plane > near_plane, "gs_depth_keys: far_plane must exceed near_plane");
  float inv_far = 0.f, inv_den = 0.f;
  if (to_ndc) {  // scalars as torch forms them: Python doubles cast to f32, the divisor inverted in f32
    inv_far = (float)(1.0 / far_plane);
    const float den = (float)(1.0 / near_plane - 1.0 / far_plane);
    inv_den = 1.0f / den;
  }
  depth_keys_kernel<<<(unsigned)ceil_div(p->num_points, 256), 256, 0, (cudaStream_t)stream>>>(
      p->num_points, p->use_depth16, to_ndc ? 1 : 0, inv_far, inv_den, depth, keys, values, count_dev);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_depth_keys(const GsTileParams* p, const float* depth, double near_plane, double far_plane, uint32_t* keys,
                  int32_t* values, void* stream) {
  return depth_keys_entry(p, depth, near_plane, far_plane, nullptr, keys, values, stream);
}

int gs_depth_keys_counted(const GsTileParams* p, const float* depth, double near_plane, double far_plane,
                          const int32_t* count_dev, uint32_t* keys, int32_t* values, void* stream) {
  GS_CHECK_ARG(count_dev != nullptr, "gs_depth_keys_counted: null count");
  return depth_keys_entry(p, depth, near_plane, far_plane, count_dev, keys, values, stream);
}

static int tile_count_perm_entry(const GsTileParams* p, const float* gaussians, const int32_t* perm,
                                 const int32_t* count_dev, int32_t* counts, uint64_t* tile_masks, void* stream) {
  int rc = check_tile_params(p, "gs_tile_count_perm");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0) return GS_OK;
  GS_CHECK_ARG(gaussians && perm && counts, "gs_tile_count_perm: null tensor");
  int ts = p->tile_size;
  int img_w = (int)ceil_div(p->image_width, ts) * ts, img_h = (int)ceil_div(p->image_height, ts) * ts;
  tile_query_kernel<0, uint32_t><<<(unsigned)ceil_div(p->num_points, kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
      *p, img_w, img_h, gaussians, nullptr, perm, nullptr, counts, nullptr, nullptr, (uint2*)tile_masks, count_dev);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_tile_count_perm(const GsTileParams* p, const float* gaussians, const int32_t* perm, int32_t* counts,
                       uint64_t* tile_masks, void* stream) {
  return tile_count_perm_entry(p, gaussians, perm, nullptr, counts, tile_masks, stream);
}

int gs_tile_count_perm_counted(const GsTileParams* p, const float* gaussians, const int32_t* perm,
                               const int32_t* count_dev, int32_t* counts, uint64_t* tile_masks, void* stream) {
  GS_CHECK_ARG(count_dev != nullptr, "gs_tile_count_perm_counted: null count");
  return tile_count_perm_entry(p, gaussians, perm, count_dev, counts, tile_masks, stream);
}

int gs_tile_emit_tiles(const GsTileParams* p, const float* gaussians, const int32_t* perm, const int32_t* cum,
                       const uint64_t* tile_masks, uint32_t* tile_ids, int32_t* values, void* stream) {
  int rc = check_tile_params(p, "gs_tile_emit_tiles");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0) return GS_OK;
  GS_CHECK_ARG(gaussians && perm && cum && tile_ids && values, "gs_tile_emit_tiles: null tensor");
  int ts = p->tile_size;
  int img_w = (int)ceil_div(p->image_width, ts) * ts, img_h = (int)ceil_div(p->image_height, ts) * ts;
  tile_query_kernel<2, uint32_t><<<(unsigned)ceil_div(p->num_points, kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
      *p, img_w, img_h, gaussians, nullptr, perm, cum, nullptr, tile_ids, values, (uint2*)tile_masks);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

static int tile_emit_tiles_capped_entry(const GsTileParams* p, const float* gaussians, const int32_t* perm,
                                        const int32_t* cum, const uint64_t* tile_masks, const int32_t* count_dev,
                                        int64_t capacity, uint32_t* tile_ids, int32_t* values, void* stream) {
  int rc = check_tile_params(p, "gs_tile_emit_tiles_capped");
  if (rc != GS_OK) return rc;
  if (p->num_points == 0 || capacity <= 0) return GS_OK;
  GS_CHECK_ARG(gaussians && perm && cum && tile_ids && values, "gs_tile_emit_tiles_capped: null tensor");
  int ts = p->tile_size;
  int img_w = (int)ceil_div(p->image_width, ts) * ts, img_h = (int)ceil_div(p->image_height, ts) * ts;
  tile_query_kernel<2, uint32_t><<<(unsigned)ceil_div(p->num_points, kTileBlock), kTileBlock, 0, (cudaStream_t)stream>>>(
      *p, img_w, img_h, gaussians, nullptr, perm, cum, nullptr, tile_ids, values, (uint2*)tile_masks, count_dev,
      (long long)capacity);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_tile_emit_tiles_capped(const GsTileParams* p, const float* gaussians, const int32_t* perm, const int32_t* cum,
                              const uint64_t* tile_masks, int64_t capacity, uint32_t* tile_ids, int32_t* values,
                              void* stream) {
  return tile_emit_tiles_capped_entry(p, gaussians, perm, cum, tile_masks, nullptr, capacity, tile_ids, values, stream);
}

int gs_tile_emit_tiles_capped_counted(const GsTileParams* p, const float* gaussians, const int32_t* perm,
                                      const int32_t* cum, const uint64_t* tile_masks, const int32_t* count_dev,
                                      int64_t capacity, uint32_t* tile_ids, int32_t* values, void* stream) {
  GS_CHECK_ARG(count_dev != nullptr, "gs_tile_emit_tiles_capped_counted: null count");
  return tile_emit_tiles_capped_entry(p, gaussians, perm, cum, tile_masks, count_dev, capacity, tile_ids, values, stream);
}

int gs_find_ranges_tiles_counted(const GsTileParams* p, int64_t capacity, const int32_t* num_overlaps_dev,
                                 const uint32_t* sorted_tile_ids, int32_t* tile_ranges, void* stream) {
  int rc = check_tile_params(p, "gs_find_ranges_tiles_counted");
  if (rc != GS_OK) return rc;
  GS_CHECK_ARG(tile_ranges != nullptr && capacity >= 0 && num_overlaps_dev != nullptr,
               "gs_find_ranges_tiles_counted: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t tiles = ceil_div(p->image_width, p->tile_size) * ceil_div(p->image_height, p->tile_size);
  GS_CUDA(cudaMemsetAsync(tile_ranges, 0, (size_t)tiles * 2 * sizeof(int32_t), st));
  if (capacity == 0) return GS_OK;
  GS_CHECK_ARG(sorted_tile_ids != nullptr, "gs_find_ranges_tiles_counted: null keys");
  find_ranges_kernel<uint32_t, 0><<<(unsigned)ceil_div(capacity, 256), 256, 0, st>>>(capacity, sorted_tile_ids, tile_ranges,
                                                                                  num_overlaps_dev);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_find_ranges_tiles(const GsTileParams* p, int64_t num_overlaps, const uint32_t* sorted_tile_ids,
                         int32_t* tile_ranges, void* stream) {
  int rc = check_tile_params(p, "gs_find_ranges_tiles");
  if (rc != GS_OK) return rc;
  GS_CHECK_ARG(tile_ranges != nullptr && num_overlaps >= 0, "gs_find_ranges_tiles: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t tiles = ceil_div(p->image_width, p->tile_size) * ceil_div(p->image_height, p->tile_size);
  GS_CUDA(cudaMemsetAsync(tile_ranges, 0, (size_t)tiles * 2 * sizeof(int32_t), st));
  if (num_overlaps == 0) return GS_OK;
  GS_CHECK_ARG(sorted_tile_ids != nullptr, "gs_find_ranges_tiles: null keys");
  find_ranges_kernel<uint32_t, 0><<<(unsigned)ceil_div(num_overlaps, 256), 256, 0, st>>>(num_overlaps, sorted_tile_ids,
                                                                                       tile_ranges);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- host image of the device math, for CPU-side op-order checks against the oracle (tests only).
// Not a compute path: evaluates ONE gaussian per call.
int gs_selftest_project_one_f32(const float* pos, const float* ls, const float* q, float logit, const float* Tcw16,
                                const float* proj4, int w, int h, double near_, double far_, double blur,
                                double margin, double thr, float* out8, int* in_view) {
  GsProjectParams p;
  p.dtype = GS_F32; p.image_width = w; p.image_height = h; p.num_points = 1; p.near_plane = near_;
  p.far_plane = far_; p.blur_cov = blur; p.clamp_margin = margin; p.alpha_threshold = thr;
  CameraConst<float> C;
  for (int i = 0; i < 12; ++i) C.Tcw[i] = Tcw16[i];
  C.fx = proj4[0]; C.fy = proj4[1]; C.cx = proj4[2]; C.cy = proj4[3];
  C.w = (float)w; C.h = (float)h; C.near_ = (float)near_; C.far_ = (float)far_; C.blur = (float)blur;
  C.lo_x = (float)(-(double)w * margin); C.lo_y = (float)(-(double)h * margin);
  C.hi_x = (float)(((double)w - 1.0) * (1.0 + margin)); C.hi_y = (float)(((double)h - 1.0) * (1.0 + margin));
  C.alpha_threshold = (float)thr;
  Projected<float> o = project_one<float>(pos, ls, q, logit, C);
  out8[0] = o.mean_x; out8[1] = o.mean_y; out8[2] = o.axis_x; out8[3] = o.axis_y;
  out8[4] = o.sigma_x; out8[5] = o.sigma_y; out8[6] = o.alpha; out8[7] = o.z;
  *in_view = o.in_view ? 1 : 0;
  return GS_OK;
}

// Overlapped tile ids of ONE gaussian in emission order; returns the count (tests only).
int gs_selftest_tile_query(const float* g7, int img_w_padded, int img_h_padded, int ts, float thr, int32_t* tile_ids,
                           int capacity) {
  TileQuery qy = obb_query(g7[0], g7[1], g7[2], g7[3], g7[4], g7[5], g7[6], img_w_padded, img_h_padded, ts, thr);
  int tiles_wide = img_w_padded / ts, n = 0;
  for (int tu = 0; tu < qy.span_x; ++tu)
    for (int tv = 0; tv < qy.span_y; ++tv)
      if (test_tile(qy, tu, tv, ts)) {
        if (tile_ids && n < capacity) tile_ids[n] = (tu + qy.min_x) + (tv + qy.min_y) * tiles_wide;
        ++n;
      }
  return n;
}

}  // extern "C"
