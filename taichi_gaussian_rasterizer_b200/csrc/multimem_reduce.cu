// multimem_reduce.cu — sum of one f32 buffer over the GPUs of an NVSwitch box, in place, reduced INSIDE the switch.
//
// The step's collective (distributed.GradientBucket.all_reduce: 708 MB of gradients at 3 M gaussians, SH degree 3).
// Every rank holds its replica of the buffer in symmetric memory with a MULTICAST mapping (one virtual address that
// stands for the same offset on every GPU; set up by torch.distributed._symmetric_memory, which is plumbing here as
// torch.distributed is).  Two-shot through the switch, one kernel:
//   1. barrier: every rank's replica is complete (flags in peer memory, release / acquire at system scope);
//   2. rank r owns the r-th slice: multimem.ld_reduce.add.v4.f32 fetches the SUM over all replicas of a 16 B vector
//      (the switch reads the replicas and adds them: 1 / world of the buffer comes down the link, not world - 1
//      partial buffers), multimem.st.v4.f32 writes it to ALL replicas (the switch fans it out);
//   3. barrier: every rank's stores have landed everywhere.
// Per GPU and direction about (1 + 1 / world) x the buffer crosses NVLink — a ring all-reduce moves 2 (world - 1) /
// world x — and no SM ever adds anything.  The reference has no collective (single process); this replaces the
// ncclAllReduce the bucket would otherwise issue.
#include "common.cuh"

namespace gs {

__device__ __forceinline__ void mm_ld_reduce(const float* mc, float4& v) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mm_ld_reduce1(const float* mc, float& v) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
}
__device__ __forceinline__ void mm_st1(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}

// flag protocol: the sender turns the peer's word 0 -> 1 (spinning while it is still 1 from the previous round), the
// receiver turns its own word 1 -> 0: self resetting, no epochs.  Word (block, sender) of every rank's flag array.
__device__ __forceinline__ void flag_put(uint32_t* addr) {
  uint32_t old;
  do {
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void flag_wait(uint32_t* addr) {
  uint32_t old;
  do {
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 1u);
}

// Block b of every rank meets block b of every other rank.  flags[q] = rank q's flag array (world x blocks words).
__device__ __forceinline__ void cross_rank_barrier(uint32_t* const* flags, int rank, int world, int phase, int nblocks,
                                                   int block) {
  // phase = 2 * channel (+ 1 for the closing barrier): reductions that may be in flight at the same time (the early
  // SH reduction under the last view and a later one) use different channels, i.e. different words
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    const size_t slot = ((size_t)phase * nblocks + block) * world;
    flag_put(flags[peer] + slot + rank);    // "I am here", written into the peer's memory
    flag_wait(flags[rank] + slot + peer);   // the peer's word in mine
  }
  __syncthreads();
}

constexpr int kMmThreads = 512;
constexpr int kMmUnroll = 4;

__global__ void __launch_bounds__(kMmThreads)
multimem_all_reduce_kernel(float* __restrict__ mc, int64_t n, int rank, int world, uint32_t* const* __restrict__ flags,
                           int channel) {
  __threadfence_system();
  cross_rank_barrier(flags, rank, world, 2 * channel, gridDim.x, blockIdx.x);
  const int64_t n4 = n / 4;
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = (int64_t)rank * per, hi = min(n4, lo + per);
  const int64_t stride = (int64_t)gridDim.x * kMmThreads;
  float4* mc4 = reinterpret_cast<float4*>(mc);
  int64_t v = lo + (int64_t)blockIdx.x * kMmThreads + threadIdx.x;
  for (; v + (kMmUnroll - 1) * stride < hi; v += kMmUnroll * stride) {
    float4 x[kMmUnroll];
#pragma unroll
    for (int u = 0; u < kMmUnroll; ++u) mm_ld_reduce(reinterpret_cast<const float*>(mc4 + v + u * stride), x[u]);
#pragma unroll
    for (int u = 0; u < kMmUnroll; ++u) mm_st(reinterpret_cast<float*>(mc4 + v + u * stride), x[u]);
  }
  for (; v < hi; v += stride) {
    float4 x;
    mm_ld_reduce(reinterpret_cast<const float*>(mc4 + v), x);
    mm_st(reinterpret_cast<float*>(mc4 + v), x);
  }
  if (rank == 0 && blockIdx.x == 0 && (int64_t)threadIdx.x < n - 4 * n4) {   // the last n mod 4 floats
    float x;
    mm_ld_reduce1(mc + 4 * n4 + threadIdx.x, x);
    mm_st1(mc + 4 * n4 + threadIdx.x, x);
  }
  __threadfence_system();
  cross_rank_barrier(flags, rank, world, 2 * channel + 1, gridDim.x, blockIdx.x);
}

// Rendezvous of the ranks on the stream, nothing else: everything every rank enqueued before it is complete (and
// visible at system scope) when it returns — e.g. the staged colour gradients the peers are about to read.
__global__ void cross_rank_barrier_kernel(int rank, int world, uint32_t* const* __restrict__ flags, int channel,
                                          int nblocks) {
  __threadfence_system();
  cross_rank_barrier(flags, rank, world, 2 * channel, nblocks, 0);   // block 0's words of that channel
}

}  // namespace gs

using namespace gs;

extern "C" {

int gs_multimem_all_reduce_flag_words(int32_t world, int32_t num_blocks, int32_t num_channels) {
  return 2 * world * num_blocks * num_channels;
}

int gs_cross_rank_barrier(int32_t rank, int32_t world, uint32_t* const* flags_dev, int32_t num_blocks, int32_t channel,
                          void* stream) {
  GS_CHECK_ARG(flags_dev != nullptr && world >= 1 && world <= 32 && rank >= 0 && rank < world && channel >= 0,
               "gs_cross_rank_barrier: bad arguments");
  // one CTA; num_blocks fixes the word layout it shares with gs_multimem_all_reduce (the slots depend on the grid size)
  GS_CHECK_ARG(num_blocks >= 1, "gs_cross_rank_barrier: bad num_blocks");
  cross_rank_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rank, world, flags_dev, channel, num_blocks);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_multimem_all_reduce(float* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world,
                           uint32_t* const* flags_dev, int32_t num_blocks, int32_t channel, void* stream) {
  GS_CHECK_ARG(channel >= 0, "gs_multimem_all_reduce: bad channel");
  GS_CHECK_ARG(multicast_ptr != nullptr && flags_dev != nullptr, "gs_multimem_all_reduce: null pointer");
  GS_CHECK_ARG(num_floats >= 0 && world >= 1 && world <= 32 && rank >= 0 && rank < world && num_blocks >= 1,
               "gs_multimem_all_reduce: bad sizes");
  GS_CHECK_ARG(((uintptr_t)multicast_ptr & 15) == 0, "gs_multimem_all_reduce: the buffer must be 16 B aligned");
  if (num_floats == 0) return GS_OK;
  multimem_all_reduce_kernel<<<num_blocks, kMmThreads, 0, (cudaStream_t)stream>>>(multicast_ptr, num_floats, rank, world,
                                                                                   flags_dev, channel);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // extern "C"
