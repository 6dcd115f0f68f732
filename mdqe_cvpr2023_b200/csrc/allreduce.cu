// Gradient all-reduce of clip-sharded data-parallel training (the reference leaves it to PyTorch DDP over NCCL,
// train_net.py:256-271) as ONE small kernel over NVLink peer memory.
//
// Why not NCCL here: the buckets are reduced while the encoder backward -- a latency-bound kernel that wants every SM -- is still
// running.  NCCL's ring needs its full set of channels (24-32 CTAs of 512+ threads, a FIFO protocol with per-step flag
// handshakes) to move 78 MB inside that window; capped to 8 / 4 / 2 CTAs the step grows from 1.98 to 2.28 / 2.93 / 4.56 ms on
// two B200s (profiles/r02_allreduce.md).  Plain bulk loads and stores over NVLink need no protocol: a handful of CTAs with deep
// memory-level parallelism saturate the links, and on an NVSwitch box the switch itself adds the replicas (multimem.ld_reduce)
// so every element crosses the SM once.
//
// Two algorithms on a buffer that every rank has mapped at `peer[r]` (symmetric allocation: same size, same offsets):
//   * multimem (NVLS): rank r owns slice r of the bucket; v = multimem.ld_reduce.add.v4.f32 [mc + i] pulls the sum over all
//     replicas out of the switch, multimem.st.v4.f32 [mc + i] broadcasts it back into every replica.
//   * two-shot P2P: rank r owns slice r; it loads the slice from every peer, adds in rank order (so every rank computes
//     bit-identical sums), and stores the result into every peer's replica.
// Both are bracketed by a cross-rank barrier per CTA pair (CTA b of every rank with CTA b of every other rank): a 0 -> 1 CAS
// on the peer's flag word (release) answered by a 1 -> 0 CAS on the own word (acquire).  The words return to 0, so the barrier
// needs no epoch and replays inside a CUDA graph.  Spins are bounded: a rank that never shows up raises an error flag instead of
// hanging the GPU.
#include <cstdint>
#include <cstdio>

#include "msda_common.cuh"
#include "msda_internal.h"

namespace msda {
namespace {

constexpr int kMaxRanks = 8;
constexpr int kArThreads = 512;       // half an SM at most (<= 64 registers would be a quarter): the kernel runs beside the backward
constexpr unsigned long long kSpinLimit = 1ull << 31;      // ~ seconds; a healthy barrier takes microseconds

struct ArArgs {
  float* peer[kMaxRanks];          // the bucket as mapped on this rank, one pointer per rank (peer[rank] = local)
  uint32_t* flags[kMaxRanks];      // flag words [n_ctas][kMaxRanks] of every rank, zero before the first use
  float* mc;                       // multicast mapping of the same buffer (multimem algorithm) or nullptr
  int rank, world;
  int* error;                      // device word on this rank: set to 1 when a spin gave up
};

__device__ __forceinline__ bool cas_spin(uint32_t* addr, uint32_t expect, uint32_t desired, bool acquire) {
  unsigned long long n = 0;
  uint32_t old;
  do {
    if (acquire)
      asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    else
      asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    if (old == expect) return true;
  } while (++n < kSpinLimit);
  return false;
}

// All ranks' CTA `blockIdx.x` meet here.  Thread t < world signals rank t and waits for rank t's signal.
__device__ __forceinline__ void cta_barrier_all_ranks(const ArArgs& a) {
  __syncthreads();                                            // every thread's earlier stores are ordered before the release below
  if (threadIdx.x < a.world) {
    const int t = threadIdx.x;
    uint32_t* theirs = a.flags[t] + blockIdx.x * kMaxRanks + a.rank;
    uint32_t* mine = a.flags[a.rank] + blockIdx.x * kMaxRanks + t;
    __threadfence_system();
    bool ok = cas_spin(theirs, 0u, 1u, false);
    ok = cas_spin(mine, 1u, 0u, true) && ok;
    if (!ok) *a.error = 1;
  }
  __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// this CTA's range of 16-byte vectors inside rank `a.rank`'s slice of [0, n4)
__device__ __forceinline__ void my_range(const ArArgs& a, int64_t n4, int64_t& begin, int64_t& end) {
  const int64_t per_rank = (n4 + a.world - 1) / a.world;
  const int64_t s0 = min(n4, per_rank * a.rank), s1 = min(n4, s0 + per_rank);
  const int64_t per_cta = (s1 - s0 + gridDim.x - 1) / gridDim.x;
  begin = min(s1, s0 + per_cta * blockIdx.x);
  end = min(s1, begin + per_cta);
}

template <int UNROLL>
__global__ void __launch_bounds__(kArThreads, 2)
allreduce_multimem_kernel(ArArgs a, int64_t off4, int64_t n4, float scale) {
  cta_barrier_all_ranks(a);                                   // every rank's bucket is final
  int64_t begin, end;
  my_range(a, n4, begin, end);
  float* mc = a.mc + 4 * off4;
  for (int64_t i = begin + threadIdx.x; i < end; i += static_cast<int64_t>(kArThreads) * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = i + static_cast<int64_t>(u) * kArThreads;
      if (j < end) v[u] = multimem_ld_reduce(mc + 4 * j);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = i + static_cast<int64_t>(u) * kArThreads;
      if (j < end) {
        v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
        multimem_st(mc + 4 * j, v[u]);
      }
    }
  }
  cta_barrier_all_ranks(a);                                   // every slice has been written back into every replica
}

template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(kArThreads, 2)
allreduce_p2p_kernel(ArArgs a, int64_t off4, int64_t n4, float scale) {
  cta_barrier_all_ranks(a);
  int64_t begin, end;
  my_range(a, n4, begin, end);
  for (int64_t i = begin + threadIdx.x; i < end; i += static_cast<int64_t>(kArThreads) * UNROLL) {
    float4 v[WORLD][UNROLL];                                  // every load of the step in flight before the first add: NVLink
#pragma unroll                                                // round trips are ~1 us, so bytes in flight are what a CTA's rate is made of
    for (int r = 0; r < WORLD; ++r) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t j = i + static_cast<int64_t>(u) * kArThreads;
        v[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < end) v[r][u] = __ldcg(reinterpret_cast<const float4*>(a.peer[r]) + off4 + j);
      }
    }
    float4 acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      acc[u] = v[0][u];
#pragma unroll
      for (int r = 1; r < WORLD; ++r) {                       // rank order: the same sum on every rank
        acc[u].x += v[r][u].x; acc[u].y += v[r][u].y; acc[u].z += v[r][u].z; acc[u].w += v[r][u].w;
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = i + static_cast<int64_t>(u) * kArThreads;
      if (j < end) {
        const float4 o = make_float4(acc[u].x * scale, acc[u].y * scale, acc[u].z * scale, acc[u].w * scale);
#pragma unroll
        for (int r = 0; r < WORLD; ++r) __stcg(reinterpret_cast<float4*>(a.peer[r]) + off4 + j, o);
      }
    }
  }
  cta_barrier_all_ranks(a);
}

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

int msda_allreduce_max_ranks(void) { return kMaxRanks; }
size_t msda_allreduce_flag_bytes(int n_ctas) { return static_cast<size_t>(n_ctas > 0 ? n_ctas : 0) * kMaxRanks * sizeof(uint32_t); }

int msda_allreduce_f32(void* stream, int algo, int rank, int world, const uint64_t* peer_ptrs, uint64_t multicast_ptr,
                       const uint64_t* flag_ptrs, void* error_word, int64_t offset_elems, int64_t n_elems, float scale, int n_ctas) {
  const char* who = "msda_allreduce_f32";
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !peer_ptrs || !flag_ptrs || !error_word)
    return fail(MSDA_ERR_INVALID_ARG, "%s: bad rank/world %d/%d or NULL table", who, rank, world);
  if (offset_elems < 0 || n_elems < 0 || (offset_elems & 3) || (n_elems & 3))
    return fail(MSDA_ERR_INVALID_ARG, "%s: offset (%lld) and count (%lld) must be multiples of 4 floats", who, (long long)offset_elems, (long long)n_elems);
  if (n_ctas < 1 || n_ctas > 148) return fail(MSDA_ERR_INVALID_ARG, "%s: n_ctas = %d", who, n_ctas);
  if (algo != 0 && algo != 1) return fail(MSDA_ERR_INVALID_ARG, "%s: algo %d (0 = two-shot P2P, 1 = multimem)", who, algo);
  if (algo == 1 && multicast_ptr == 0) return fail(MSDA_ERR_UNSUPPORTED, "%s: the multimem algorithm needs a multicast mapping", who);
  if (n_elems == 0) return 0;
  ArArgs a{};
  for (int r = 0; r < world; ++r) {
    if (!peer_ptrs[r] || !flag_ptrs[r] || (peer_ptrs[r] & 15u)) return fail(MSDA_ERR_INVALID_ARG, "%s: peer %d: NULL or unaligned mapping", who, r);
    a.peer[r] = reinterpret_cast<float*>(peer_ptrs[r]);
    a.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
  }
  a.mc = reinterpret_cast<float*>(multicast_ptr);
  a.rank = rank;
  a.world = world;
  a.error = static_cast<int*>(error_word);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t off4 = offset_elems / 4, n4 = n_elems / 4;
  if (algo == 1) {
    allreduce_multimem_kernel<8><<<n_ctas, kArThreads, 0, st>>>(a, off4, n4, scale);
  } else {
    switch (world) {
#define MSDA_AR_CASE(W, U) case W: allreduce_p2p_kernel<W, U><<<n_ctas, kArThreads, 0, st>>>(a, off4, n4, scale); break;
      MSDA_AR_CASE(1, 8) MSDA_AR_CASE(2, 4) MSDA_AR_CASE(3, 3) MSDA_AR_CASE(4, 2)
      MSDA_AR_CASE(5, 2) MSDA_AR_CASE(6, 1) MSDA_AR_CASE(7, 1) MSDA_AR_CASE(8, 1)
#undef MSDA_AR_CASE
    }
  }
  return after_launch(algo == 1 ? "allreduce_multimem_kernel" : "allreduce_p2p_kernel");
}

}  // extern "C"
