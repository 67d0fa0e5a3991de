// lookback.cuh — single-pass chained scan support (decoupled look-back).
//
// A status word per tile packs {flag:2 | value:62}; flag 1 = tile aggregate published, 2 = inclusive
// prefix published.  Tiles take a ticket from an atomic counter when they start, so every
// predecessor of a tile is already resident or finished and the spin cannot deadlock.
#pragma once

#include "common.cuh"

namespace gs {

constexpr unsigned long long kFlagAggregate = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by all 32 lanes of one warp of the tile.  Returns the exclusive prefix of `aggregate`
// over all earlier tiles (valid in every lane).
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status, int tile,
                                                                 unsigned long long aggregate) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_status(status, kFlagPrefix | aggregate);
    return 0ull;
  }
  if (lane == 0) st_status(status + tile, kFlagAggregate | aggregate);
  unsigned long long exclusive = 0;
  int j = tile - 1;
  while (true) {
    int src = j - lane;
    unsigned long long s;
    do {
      s = (src >= 0) ? ld_status(status + src) : kFlagPrefix;
    } while (__any_sync(kFull, (s >> 62) == 0));
    unsigned pm = __ballot_sync(kFull, (s >> 62) == 2);
    int first = pm ? (__ffs(pm) - 1) : 32;
    unsigned long long v = (lane <= first) ? (s & kValueMask) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    exclusive += v;
    if (pm) break;
    j -= 32;
  }
  if (lane == 0) st_status(status + tile, kFlagPrefix | ((exclusive + aggregate) & kValueMask));
  return exclusive;
}

}  // namespace gs
