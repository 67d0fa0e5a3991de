// scan_sort.cu — full_cumsum (single-pass chained scan) and a stable onesweep LSD radix sort of
// (key, int32 value) pairs.
//
// Replaces the reference's CUB wrappers (paths relative to /root/reference/taichi_splatting/):
//   cuda_lib/full_cumsum.cu:16-67        cub::DeviceScan::ExclusiveSum + complete_cumsum + device sync
//   cuda_lib/radix_sort_pairs.cu:9-69    cub::DeviceRadixSort::SortPairs + device sync
// Both are integer / byte work bound by HBM: scan moves 8 B per element, the sort 8 B per key for the
// histogram plus (2*key + 2*4) B per key per 8-bit pass.  No host synchronisation happens here.
#include "common.cuh"
#include "lookback.cuh"

namespace gs {

// ------------------------------------------------------------------------------------------------ scan
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

template <typename T>
__global__ void __launch_bounds__(kScanBlock)
full_cumsum_kernel(int64_t n, const T* __restrict__ in, T* __restrict__ out, unsigned long long* status,
                   unsigned int* ticket, const int32_t* __restrict__ n_dev = nullptr, T* __restrict__ total_out = nullptr) {
  // counted variant: the grid covers the capacity n, only the first *n_dev elements count (the rest scan as zeros)
  // and the total also goes to total_out, a fixed address the host can read without knowing the count
  if (n_dev) n = min(n, (int64_t)*n_dev);
  __shared__ int s_tile;
  __shared__ unsigned long long s_warp_total[kScanBlock / 32];
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;

  T v[kScanItems];
  unsigned long long thread_sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : T(0);
    thread_sum += (unsigned long long)v[k];
  }
  unsigned long long incl = thread_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_total[warp] = incl;
  __syncthreads();
  unsigned long long warp_offset = 0, block_total = 0;
#pragma unroll
  for (int wi = 0; wi < kScanBlock / 32; ++wi) {
    unsigned long long c = s_warp_total[wi];
    if (wi < warp) warp_offset += c;
    block_total += c;
  }
  if (warp == 0) {
    unsigned long long ex = lookback_exclusive(status, tile, block_total);
    if (lane == 0) {
      s_prefix = ex;
      if (tile == (int)gridDim.x - 1) {
        out[n] = (T)(ex + block_total);
        if (total_out) *total_out = (T)(ex + block_total);
      }
    }
  }
  __syncthreads();
  unsigned long long run = s_prefix + warp_offset + (incl - thread_sum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = (T)run;
    run += (unsigned long long)v[k];
  }
}

// ------------------------------------------------------------------------------------------------ radix sort
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortBlock = 256;  // == kRadix: thread t owns digit t in the per-digit phases
constexpr int kSortWarps = kSortBlock / 32;
constexpr int kMaxPasses = 8;
constexpr unsigned kSortFlagAgg = 1u << 30, kSortFlagPrefix = 2u << 30, kSortValueMask = (1u << 30) - 1u;

template <typename KeyT> struct SortCfg;
template <> struct SortCfg<uint32_t> { static constexpr int kItems = 16; };
template <> struct SortCfg<uint64_t> { static constexpr int kItems = 16; };

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Digit histograms of every pass in one sweep over the keys (8 B/key of HBM traffic).
template <typename KeyT>
__global__ void __launch_bounds__(256)
radix_histogram_kernel(int64_t n, const KeyT* __restrict__ keys, int begin_bit, int end_bit, int passes,
                       unsigned* __restrict__ global_hist, const int32_t* __restrict__ n_dev) {
  __shared__ unsigned h[kMaxPasses * kRadix];
  if (n_dev) n = min(n, (int64_t)*n_dev);   // counted variant: n is the capacity, the count is on the device
  for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const KeyT k = keys[i];
#pragma unroll 1
    for (int p = 0; p < passes; ++p) {
      const int shift = begin_bit + p * kRadixBits;
      const int nb = min(kRadixBits, end_bit - shift);
      atomicAdd(&h[p * kRadix + (unsigned)((k >> shift) & (KeyT)((1u << nb) - 1u))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x)
    if (h[i]) atomicAdd(&global_hist[i], h[i]);
}

// One onesweep pass: rank keys of a tile by the current digit (warp match-any multi-split + per-warp
// shared histograms), chain the per-digit tile counts through decoupled look-back, reorder the tile in
// shared memory and write digit runs back coalesced.  Stable: ranks follow input order.
// FULL: the digit has all kRadixBits bits (every pass but a short last one): no per bit "is this bit in the mask" test.
template <typename KeyT, bool FULL>
__global__ void __launch_bounds__(kSortBlock, 3)
onesweep_pass_kernel(int64_t n, const KeyT* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                     KeyT* __restrict__ keys_out, int32_t* __restrict__ vals_out, int shift, unsigned mask,
                     const unsigned* __restrict__ global_hist, unsigned* __restrict__ status,
                     unsigned* __restrict__ ticket, const int32_t* __restrict__ n_dev) {
  constexpr int ITEMS = SortCfg<KeyT>::kItems;
  if (n_dev) n = min(n, (int64_t)*n_dev);   // counted variant: the grid covers the capacity
  constexpr int TILE = kSortBlock * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);
  int32_t* s_vals = reinterpret_cast<int32_t*>(s_keys + TILE);
  unsigned* s_warp_hist = reinterpret_cast<unsigned*>(s_vals + TILE);  // [kSortWarps][kRadix]
  unsigned* s_digit_off = s_warp_hist + kSortWarps * kRadix;           // [kRadix] exclusive offset in tile
  unsigned* s_global = s_digit_off + kRadix;                           // [kRadix] global base of the digit
  __shared__ int s_tile;
  __shared__ unsigned s_scan_warp[kSortWarps], s_scan_warp_g[kSortWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  for (int i = tid; i < kSortWarps * kRadix; i += kSortBlock) s_warp_hist[i] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int64_t tile_base = (int64_t)tile * TILE;
  if (tile_base >= n) return;   // (counted) tiles past the end: no items, and no later tile looks back at them
  const int valid = (int)min((int64_t)TILE, n - tile_base);
  const int64_t warp_base = tile_base + (int64_t)warp * (32 * ITEMS);
  // this pass's global digit counts (raw, from radix_histogram_kernel): every block scans the 256 of them itself,
  // together with its own tile counts, instead of a separate scan launch per sort
  const unsigned gcount = global_hist[tid];

  KeyT key[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int64_t idx = warp_base + k * 32 + lane;
    key[k] = idx < n ? keys_in[idx] : ~KeyT(0);
  }

  unsigned rank[ITEMS];
  unsigned* my_hist = s_warp_hist + warp * kRadix;
  const unsigned lanemask_le = 0xffffffffu >> (31 - lane);
  // Peer masks (lanes holding the same digit) from one ballot per digit bit instead of match.any: the ballots of the
  // ITEMS keys are independent and pipeline, MATCH.ANY serialises (measured 0.206 -> 0.177 ms on 3 M 32 bit pairs).
  unsigned peers[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const unsigned d = (unsigned)(key[k] >> shift) & mask;
    unsigned pm = kFull;
#pragma unroll
    for (int b = 0; b < kRadixBits; ++b) {
      if (FULL || ((mask >> b) & 1u)) {  // uniform: short last digits skip their absent bits
        const bool bit = (d & (1u << b)) != 0u;
        const unsigned bal = __ballot_sync(kFull, bit);
        pm &= bal ^ (bit ? 0u : ~0u);
      }
    }
    peers[k] = pm;
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const unsigned d = (unsigned)(key[k] >> shift) & mask;
    const int leader = 31 - __clz(peers[k]);
    const unsigned cnt_le = __popc(peers[k] & lanemask_le);
    unsigned old = 0;
    if (lane == leader) {
      old = my_hist[d];
      my_hist[d] = old + cnt_le;
    }
    __syncwarp();
    old = __shfl_sync(kFull, old, leader);
    rank[k] = old + cnt_le - 1u;
  }
  __syncthreads();

  // thread tid owns digit tid: exclusive scan of the warp histograms, then the tile count look-back
  unsigned tile_count = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    const unsigned c = s_warp_hist[w * kRadix + tid];
    s_warp_hist[w * kRadix + tid] = tile_count;
    tile_count += c;
  }
  unsigned exclusive = 0;
  {
    unsigned* st = status + (size_t)tile * kRadix + tid;
    if (tile == 0) {
      st_relaxed_u32(st, kSortFlagPrefix | tile_count);
    } else {
      st_relaxed_u32(st, kSortFlagAgg | tile_count);
      int j = tile - 1;
      // LB independent loads in flight per step: the walk is a chain of L2 round trips otherwise
      constexpr int LB = 8;
      while (true) {
        unsigned v[LB];
#pragma unroll
        for (int u = 0; u < LB; ++u)
          v[u] = j - u >= 0 ? ld_relaxed_u32(status + (size_t)(j - u) * kRadix + tid) : kSortFlagPrefix;
        bool done = false;
        int u = 0;
#pragma unroll
        for (; u < LB; ++u) {
          if ((v[u] >> 30) == 0u) break;  // not published yet: retry from here
          exclusive += v[u] & kSortValueMask;
          if ((v[u] >> 30) == 2u) { done = true; break; }
        }
        if (done) break;
        j -= u;
      }
      st_relaxed_u32(st, kSortFlagPrefix | ((exclusive + tile_count) & kSortValueMask));
    }
  }
  // exclusive scans over the 256 digits: of tile_count (offsets inside the tile) and of the global counts
  {
    unsigned incl = tile_count, gincl = gcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(kFull, incl, o), g = __shfl_up_sync(kFull, gincl, o);
      if (lane >= o) { incl += t; gincl += g; }
    }
    if (lane == 31) { s_scan_warp[warp] = incl; s_scan_warp_g[warp] = gincl; }
    __syncthreads();
    unsigned off = 0, goff = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
      if (w < warp) { off += s_scan_warp[w]; goff += s_scan_warp_g[w]; }
    s_digit_off[tid] = off + incl - tile_count;
    s_global[tid] = (goff + gincl - gcount) + exclusive;
  }
  __syncthreads();

#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const unsigned d = (unsigned)(key[k] >> shift) & mask;
    const unsigned pos = s_digit_off[d] + my_hist[d] + rank[k];
    const int64_t idx = warp_base + k * 32 + lane;
    s_keys[pos] = key[k];
    s_vals[pos] = idx < n ? vals_in[idx] : 0;
  }
  __syncthreads();

  for (int i = tid; i < valid; i += kSortBlock) {
    const KeyT k = s_keys[i];
    const unsigned d = (unsigned)(k >> shift) & mask;
    const int64_t dst = (int64_t)s_global[d] + (i - (int)s_digit_off[d]);
    keys_out[dst] = k;
    vals_out[dst] = s_vals[i];
  }
}

template <typename KeyT>
static size_t sort_smem_bytes() {
  constexpr int TILE = kSortBlock * SortCfg<KeyT>::kItems;
  return (size_t)TILE * (sizeof(KeyT) + sizeof(int32_t)) + (size_t)(kSortWarps + 2) * kRadix * sizeof(unsigned);
}

struct SortLayout {
  int passes;
  int64_t tiles;
  size_t off_hist, off_ticket, off_status, off_tmp_keys, off_tmp_vals, total, zero_bytes;
};

static SortLayout sort_layout(int64_t n, int key_bytes, int begin_bit, int end_bit) {
  SortLayout L;
  L.passes = (int)ceil_div(end_bit - begin_bit, kRadixBits);
  int items = key_bytes == 8 ? SortCfg<uint64_t>::kItems : SortCfg<uint32_t>::kItems;
  L.tiles = ceil_div(n > 0 ? n : 1, (int64_t)kSortBlock * items);
  size_t off = 0;
  L.off_hist = off; off += (size_t)kMaxPasses * kRadix * sizeof(unsigned);
  L.off_ticket = off; off += 64;
  L.off_status = off; off += (size_t)L.passes * L.tiles * kRadix * sizeof(unsigned);
  L.zero_bytes = off;  // everything up to here must be zero before the first kernel
  off = align_up(off, 256);
  L.off_tmp_keys = off; off += align_up((size_t)n * key_bytes, 256);
  L.off_tmp_vals = off; off += align_up((size_t)n * sizeof(int32_t), 256);
  L.total = off;
  return L;
}

template <typename KeyT>
static int radix_sort_impl(int64_t n, const KeyT* keys_in, const int32_t* vals_in, KeyT* keys_out, int32_t* vals_out,
                           int begin_bit, int end_bit, unsigned char* ws, cudaStream_t st,
                           const int32_t* n_dev = nullptr) {
  const SortLayout L = sort_layout(n, sizeof(KeyT), begin_bit, end_bit);
  unsigned* hist = (unsigned*)(ws + L.off_hist);
  unsigned* tickets = (unsigned*)(ws + L.off_ticket);
  unsigned* status = (unsigned*)(ws + L.off_status);
  KeyT* tmp_keys = (KeyT*)(ws + L.off_tmp_keys);
  int32_t* tmp_vals = (int32_t*)(ws + L.off_tmp_vals);

  GS_CUDA(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
  int hist_blocks = (int)min((int64_t)148 * 8, ceil_div(n, 256));
  radix_histogram_kernel<KeyT><<<hist_blocks, 256, 0, st>>>(n, keys_in, begin_bit, end_bit, L.passes, hist, n_dev);
  GS_LAUNCH_CHECK();

  const size_t smem = sort_smem_bytes<KeyT>();
  GS_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GS_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const KeyT* src_k = keys_in;
  const int32_t* src_v = vals_in;
  for (int p = 0; p < L.passes; ++p) {
    // alternate between the temporary and the output so that the last pass lands in the output
    const bool to_out = ((L.passes - 1 - p) % 2) == 0;
    KeyT* dst_k = to_out ? keys_out : tmp_keys;
    int32_t* dst_v = to_out ? vals_out : tmp_vals;
    const int shift = begin_bit + p * kRadixBits;
    const int nb = min(kRadixBits, end_bit - shift);
    auto kern = nb == kRadixBits ? onesweep_pass_kernel<KeyT, true> : onesweep_pass_kernel<KeyT, false>;
    kern<<<(unsigned)L.tiles, kSortBlock, smem, st>>>(
        n, src_k, src_v, dst_k, dst_v, shift, (1u << nb) - 1u, hist + p * kRadix,
        status + (size_t)p * L.tiles * kRadix, tickets + p, n_dev);
    GS_LAUNCH_CHECK();
    src_k = dst_k;
    src_v = dst_v;
  }
  return GS_OK;
}

}  // namespace gs

using namespace gs;

extern "C" {

size_t gs_full_cumsum_workspace_bytes(int64_t n, int32_t /*elem_bytes*/) {
  int64_t tiles = ceil_div(n > 0 ? n : 1, kScanTile);
  return align_up((size_t)tiles * sizeof(unsigned long long) + 16, 256);
}

static int full_cumsum_entry(int64_t n, int32_t elem_bytes, const int32_t* count_dev, const void* in, void* out,
                             void* total_out, void* workspace, size_t workspace_bytes, void* stream) {
  GS_CHECK_ARG(n >= 0 && out != nullptr, "gs_full_cumsum: bad arguments");
  GS_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "gs_full_cumsum: elem_bytes must be 4 or 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    GS_CUDA(cudaMemsetAsync(out, 0, elem_bytes, st));
    if (total_out) GS_CUDA(cudaMemsetAsync(total_out, 0, elem_bytes, st));
    return GS_OK;
  }
  GS_CHECK_ARG(in != nullptr, "gs_full_cumsum: null input");
  size_t need = gs_full_cumsum_workspace_bytes(n, elem_bytes);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gs_full_cumsum: workspace %zu < %zu bytes", workspace_bytes, need);
    return GS_ERR_WORKSPACE;
  }
  int64_t tiles = ceil_div(n, kScanTile);
  GS_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  unsigned long long* status = (unsigned long long*)workspace;
  unsigned int* ticket = (unsigned int*)(status + tiles);
  if (elem_bytes == 4)
    full_cumsum_kernel<int32_t><<<(unsigned)tiles, kScanBlock, 0, st>>>(n, (const int32_t*)in, (int32_t*)out, status,
                                                                       ticket, count_dev, (int32_t*)total_out);
  else
    full_cumsum_kernel<long long><<<(unsigned)tiles, kScanBlock, 0, st>>>(n, (const long long*)in, (long long*)out,
                                                                         status, ticket, count_dev,
                                                                         (long long*)total_out);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_full_cumsum(int64_t n, int32_t elem_bytes, const void* in, void* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  return full_cumsum_entry(n, elem_bytes, nullptr, in, out, nullptr, workspace, workspace_bytes, stream);
}

int gs_full_cumsum_counted(int64_t capacity, int32_t elem_bytes, const int32_t* count_dev, const void* in, void* out,
                           void* total_out, void* workspace, size_t workspace_bytes, void* stream) {
  GS_CHECK_ARG(count_dev != nullptr && total_out != nullptr, "gs_full_cumsum_counted: null count / total");
  return full_cumsum_entry(capacity, elem_bytes, count_dev, in, out, total_out, workspace, workspace_bytes, stream);
}

size_t gs_radix_sort_pairs_workspace_bytes(int64_t n, int32_t key_bytes, int32_t begin_bit, int32_t end_bit) {
  if (n < 0 || (key_bytes != 4 && key_bytes != 8) || end_bit <= begin_bit) return 0;
  return sort_layout(n, key_bytes, begin_bit, end_bit).total;
}

static int radix_sort_entry(int64_t n, const int32_t* count_dev, int32_t key_bytes, const void* keys_in,
                            const int32_t* values_in, void* keys_out, int32_t* values_out, int32_t begin_bit,
                            int32_t end_bit, void* workspace, size_t workspace_bytes, void* stream) {
  GS_CHECK_ARG(key_bytes == 4 || key_bytes == 8, "gs_radix_sort_pairs: key_bytes must be 4 or 8");
  GS_CHECK_ARG(begin_bit >= 0 && end_bit > begin_bit && end_bit <= key_bytes * 8,
               "gs_radix_sort_pairs: bad bit range [%d, %d)", begin_bit, end_bit);
  GS_CHECK_ARG(n >= 0 && n < (1ll << 30), "gs_radix_sort_pairs: n must be in [0, 2^30)");
  if (n == 0) return GS_OK;
  GS_CHECK_ARG(keys_in && values_in && keys_out && values_out, "gs_radix_sort_pairs: null tensor");
  GS_CHECK_ARG(ceil_div(end_bit - begin_bit, kRadixBits) <= kMaxPasses, "gs_radix_sort_pairs: too many passes");
  size_t need = gs_radix_sort_pairs_workspace_bytes(n, key_bytes, begin_bit, end_bit);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gs_radix_sort_pairs: workspace %zu < %zu bytes", workspace_bytes, need);
    return GS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (key_bytes == 8)
    return radix_sort_impl<uint64_t>(n, (const uint64_t*)keys_in, values_in, (uint64_t*)keys_out, values_out,
                                     begin_bit, end_bit, (unsigned char*)workspace, st, count_dev);
  return radix_sort_impl<uint32_t>(n, (const uint32_t*)keys_in, values_in, (uint32_t*)keys_out, values_out, begin_bit,
                                   end_bit, (unsigned char*)workspace, st, count_dev);
}

int gs_radix_sort_pairs(int64_t n, int32_t key_bytes, const void* keys_in, const int32_t* values_in, void* keys_out,
                        int32_t* values_out, int32_t begin_bit, int32_t end_bit, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return radix_sort_entry(n, nullptr, key_bytes, keys_in, values_in, keys_out, values_out, begin_bit, end_bit,
                          workspace, workspace_bytes, stream);
}

int gs_radix_sort_pairs_counted(int64_t capacity, const int32_t* count_dev, int32_t key_bytes, const void* keys_in,
                                const int32_t* values_in, void* keys_out, int32_t* values_out, int32_t begin_bit,
                                int32_t end_bit, void* workspace, size_t workspace_bytes, void* stream) {
  GS_CHECK_ARG(count_dev != nullptr, "gs_radix_sort_pairs_counted: null count");
  return radix_sort_entry(capacity, count_dev, key_bytes, keys_in, values_in, keys_out, values_out, begin_bit, end_bit,
                          workspace, workspace_bytes, stream);
}

}  // extern "C"
