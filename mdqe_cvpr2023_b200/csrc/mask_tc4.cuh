// 3xTF32 mask contraction, fourth generation: MN-major fp32 operands go from TMA straight into the MMA.
//
//   out[b, r, n] = sum_k A[b, r, k] * P[b, k, n]          forward : A = coeff [Q][K] (K-major), P = proto
//   out[b, r, n] = sum_k A[b, k, r] * P[b, k, n]          kTransB : A = coeff read as [q][k] with rows = k (grad_proto, P = grad_out)
//
// The plane operand P has n contiguous, i.e. it is MN-major for an MMA whose M dimension is n.  For 32-bit types the tensor
// core accepts MN-major tiles only in the "128B swizzle, 32B atom" layout (UMMA layout type 1; TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32 fp32 along n, 32-byte chunk index XOR (row & 3), atoms of 4 reduction rows.
// With that layout the third-generation kernel's on-chip transposition disappears: the split warps only do the element-wise
// hi/lo split (hi = 19-bit truncation in place, lo = a - hi into the twin tile), whatever the swizzle.
//   smem tile P  : 4 boxes {32 n, 32 k} of 4 KB; descriptor LBO = 4096 (next 32 n), SBO = 512 (next 4 k), +1024 B per MMA k-step (8)
//   smem tile A  : forward  K-major SW128 rows (q) of 32 k            (LBO unused, SBO = 1024, +32 B per k-step)
//                  kTransB  boxes {32 r, 32 k} like P                  (LBO = 4096, SBO = 512, +1024 B per k-step)
// Everything else (persistent CTA, accumulators double-buffered in TMEM, TMA bulk-store epilogue) is the tc2/tc3 design.
#pragma once

#include "mask_tc.cuh"

namespace msda {

constexpr int kTc4SplitWarps = 8;
constexpr int kTc4Threads = (2 + kTc2EpiWarps + kTc4SplitWarps) * 32;     // 576
constexpr int kTc4MaxStages = 4;

// UMMA shared-memory descriptor with an explicit layout type (1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// kind::tf32, D = fp32, M = 128, N = n; a_mn / b_mn select MN-major operands
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t n, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- A operand through tensor memory (also used by gemm3x.cuh): the staging warps write hi / lo columns with tcgen05.st and the
// MMA takes A from there ([a_tmem] form), so A costs shared memory one read instead of read + write + one fetch per MMA.
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// 16 consecutive TMEM columns of the calling warp's 32 lanes (shape 32x32b: thread i owns lane i of the warp's quarter)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__device__ __forceinline__ uint32_t tf32_lo(uint32_t v) { return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xffffe000u)); }

// TMEM column of the A chunk of pipeline stage s in the mask kernels: the two accumulators own columns [0, 128) and [256, 384)
// (QN <= 128), stages 0-1 use [128, 256), stages 2-3 [384, 512); hi at +0, lo at +32
__device__ __forceinline__ uint32_t mask_tc4_a_col(int s) { return (s < 2 ? 128u : 256u) + 64u * static_cast<uint32_t>(s); }

template <typename OT, bool kTransB, bool a_tm>
__global__ void __launch_bounds__(kTc4Threads, 1)
mask_fwd_tc4_kernel(const __grid_constant__ CUtensorMap map_plane, const __grid_constant__ CUtensorMap map_rows,
                    const __grid_constant__ CUtensorMap map_out, int Q, int n_kchunks, int QS, int QN, int n_qchunks,
                    int n_tiles_n, int n_items, int n_stages, int keep_raw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  constexpr uint32_t a_bytes = kTcTileN * 128u;                          // plane tile: 4 boxes {32 n, 32 k}
  const uint32_t b_rows = kTransB ? static_cast<uint32_t>((QN + 31) / 32 * 32) : static_cast<uint32_t>(QN);
  const uint32_t b_bytes = (b_rows * 128u + 1023u) & ~1023u;
  // a_tm (compile time: as a run-time flag the extra branch per MMA slowed the single issuing thread -- grad_proto 26 -> 36 us):
  // the plane operand goes through tensor memory (hi / lo written there by the staging warps), its lo twin in shared memory
  // disappears and the row operand moves up: [P raw][A hi][A lo] instead of [P hi][P lo][A hi][A lo]
  const uint32_t b_off = a_tm ? a_bytes : 2 * a_bytes;
  const uint32_t stage_bytes = b_off + 2 * b_bytes;
  constexpr uint32_t kOutBuf = 32u * kTcTileN * sizeof(OT);
  uint8_t* out_stage = smem + n_stages * stage_bytes;
  __shared__ __align__(8) uint64_t bars[3 * kTc4MaxStages + 4];
  __shared__ uint32_t s_tmem_base;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (kTc4MaxStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * kTc4MaxStages + s); };
  auto bar_tfull = [&](int a) { return bar0 + 8u * (3 * kTc4MaxStages + a); };
  auto bar_tempty = [&](int a) { return bar0 + 8u * (3 * kTc4MaxStages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), kTc4SplitWarps); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), kTc2EpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_plane) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rows) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  const int per_chunk = n_items / n_qchunks;
  const float inv_per_chunk = 1.f / static_cast<float>(per_chunk), inv_tiles_n = 1.f / static_cast<float>(n_tiles_n);
  const bool small_items = n_items < (1 << 23);
  auto decode = [&](int item, int& b, int& tile, int& qc) {
    if (small_items) {
      qc = fast_div_small(item, per_chunk, inv_per_chunk);
      const int t = item - qc * per_chunk;
      b = fast_div_small(t, n_tiles_n, inv_tiles_n);
      tile = t - b * n_tiles_n;
    } else {
      qc = item / per_chunk;
      const int t = item - qc * per_chunk;
      tile = t % n_tiles_n;
      b = t / n_tiles_n;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x)
        for (int kc = 0; kc < n_kchunks; ++kc, ++i) {
          const int s = i % n_stages;
          const uint32_t ph = (i / n_stages) & 1;
          int b, tile, qc;
          decode(item, b, tile, qc);
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t dst = smem_u32(smem) + s * stage_bytes;
          mbar_expect_tx(bar_full(s), a_bytes + b_rows * 128u);
          for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * 4096u, &map_plane, bar_full(s), tile * kTcTileN + j * 32, kc * 32, b);
          if constexpr (kTransB) {
            for (uint32_t j = 0; j < b_rows / 32; ++j)
              tma_load_3d(dst + b_off + j * 4096u, &map_rows, bar_full(s), qc * QS + j * 32, kc * 32, b);
          } else {
            tma_load_3d(dst + b_off, &map_rows, bar_full(s), kc * 32, qc * QS, b);
          }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(static_cast<uint32_t>(QN), a_tm ? 0u : 1u, kTransB ? 1u : 0u);
      int i = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(bar_tempty(a), aph ^ 1);
        uint32_t acc = 0;
        for (int kc = 0; kc < n_kchunks; ++kc, ++i) {
          const int s = i % n_stages;
          const uint32_t ph = (i / n_stages) & 1;
          mbar_wait(bar_ready(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem) + s * stage_bytes, a_lo = a_hi + a_bytes;
          const uint32_t b_hi = a_hi + b_off, b_lo = b_hi + b_bytes;
          const uint32_t a_sel[3] = {a_hi, a_hi, a_lo}, b_sel[3] = {b_hi, b_lo, b_hi};     // hi*hi + hi*lo + lo*hi
          const uint32_t a_col = tmem_base + mask_tc4_a_col(s);
          for (int term = 0; term < 3; ++term)
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t b_desc = kTransB ? umma_desc(b_sel[term] + ks * 1024u, 4096u, 512u, 1u)
                                              : umma_desc(b_sel[term] + ks * 32u, 16u, 1024u, 2u);
              if constexpr (a_tm) {
                umma_tf32_ta(tmem_base + a * 256u, a_col + (term == 2 ? 32u : 0u) + ks * 8u, b_desc, idesc, acc);
              } else {
                const uint64_t a_desc = umma_desc(a_sel[term] + ks * 1024u, 4096u, 512u, 1u);
                umma_tf32(tmem_base + a * 256u, a_desc, b_desc, idesc, acc);
              }
              acc = 1;
            }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty(s)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull(a)) : "memory");
      }
    }
  } else if (warp >= 2 + kTc2EpiWarps) {
    // ---- split warps: element-wise, so the operand layouts do not matter
    const uint32_t t = threadIdx.x - (2 + kTc2EpiWarps) * 32;           // 0 .. 255
    const uint32_t a_vecs = a_bytes / 16, b_vecs = b_rows * 8u;
    int i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      for (int kc = 0; kc < n_kchunks; ++kc, ++i) {
        const int s = i % n_stages;
        const uint32_t ph = (i / n_stages) & 1;
        mbar_wait(bar_full(s), ph);
        uint8_t* st = smem + s * stage_bytes;
        uint4* a_hi = reinterpret_cast<uint4*>(st);
        uint4* a_lo = reinterpret_cast<uint4*>(st + a_bytes);
        uint4* b_hi = reinterpret_cast<uint4*>(st + b_off);
        uint4* b_lo = reinterpret_cast<uint4*>(st + b_off + b_bytes);
        auto split = [](uint4& v, uint4& lo) {
          uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
          uint32_t* pl = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t hi = pv[e] & 0xffffe000u;
            pl[e] = __float_as_uint(__uint_as_float(pv[e]) - __uint_as_float(hi));
            pv[e] = hi;
          }
        };
#pragma unroll
        // keep_raw: the MMA truncates fp32 to TF32 itself, so the tile as loaded is the "hi" operand (mask_tc_bwd.cuh)
        if constexpr (a_tm) {
          // plane boxes {32 n, 32 k}: row k = 128 bytes holding 32 n, 32-byte chunk index XOR (k & 3); this thread owns column
          // n = 32 * quarter + lane of the tile (= TMEM lane) and 16 of the chunk's 32 k
          const int quarter = warp & 3, khalf = (warp - (2 + kTc2EpiWarps)) >> 2;
          const uint8_t* box = st + quarter * 4096;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = khalf * 16 + j;
            hi[j] = *reinterpret_cast<const uint32_t*>(box + k * 128 + ((((lane >> 3) ^ (k & 3)) << 5) | ((lane & 7) << 2)));
            lo[j] = tf32_lo(hi[j]);
          }
          const uint32_t a_col = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + mask_tc4_a_col(s) + khalf * 16u;
          tmem_st16(a_col, hi);
          tmem_st16(a_col + 32u, lo);
        } else {
          for (uint32_t k = t; k < a_vecs; k += kTc4SplitWarps * 32) { uint4 v = a_hi[k], lo; split(v, lo); if (!keep_raw) a_hi[k] = v; a_lo[k] = lo; }
        }
        for (uint32_t k = t; k < b_vecs; k += kTc4SplitWarps * 32) { uint4 v = b_hi[k], lo; split(v, lo); if (!keep_raw) b_hi[k] = v; b_lo[k] = lo; }
        if constexpr (a_tm) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if constexpr (a_tm) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(s));
      }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const bool is_issuer = ((warp - 2) & 3) == 0 && lane == 0;
    OT* my_stage = reinterpret_cast<OT*>(out_stage + (half * 2) * kOutBuf);
    uint32_t use = 0;
    int i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
      const int a = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      int b, tile, qc;
      decode(item, b, tile, qc);
      const int q_begin = qc * QS;
      const int rows = (qc == n_qchunks - 1) ? (Q - q_begin) : QS;
      mbar_wait(bar_tfull(a), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = tmem_base + a * 256u + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int q0 = half * 32; q0 < rows; q0 += 64, ++use) {
        OT* buf = my_stage + (use & 1) * (32 * kTcTileN);
        if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1 + half, 128);
        float v[32];
        tmem_ld32(lane_base + static_cast<uint32_t>(q0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) st_stage(buf + j * kTcTileN + quarter * 32 + lane, v[j]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(1 + half, 128);
        if (is_issuer) tma_store_3d(&map_out, smem_u32(buf), tile * kTcTileN, q_begin + q0, b);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(a));
    }
    if (is_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

inline size_t mask_tc4_stage_bytes(int QN, bool trans_b, bool a_tm) {
  const size_t b_rows = trans_b ? static_cast<size_t>((QN + 31) / 32 * 32) : static_cast<size_t>(QN);
  const size_t b_bytes = (b_rows * 128 + 1023) & ~size_t(1023);
  return (a_tm ? 1 : 2) * static_cast<size_t>(kTcTileN) * 128 + 2 * b_bytes;
}

}  // namespace msda
