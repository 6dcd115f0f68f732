// Tensor-core mask contraction for sm_100a: tcgen05.mma with the accumulator in TMEM, operands staged
// by TMA (cp.async.bulk.tensor) into 128B-swizzled shared memory.
//
//   out[b, q, n] = sum_k coeff[b, q, k] * proto[b, k, n]          (einsum 'bqm,bmthw->bqthw',
//   /root/reference/mdqe/models/matcher.py:182, criterion.py:440, transformer_dec.py:255, mdqe/mdqe.py:384)
//
// Shape of the problem: K = hidden_dim/8 = 32 (24 for Swin-L), Q ~ 200, N = T*H/4*W/4 ~ 6e4..2e5.  The
// contraction is bound by WRITING the Q x N result (arithmetic intensity ~14-28 flop/B), so the design
// goal is an epilogue that streams TMEM -> registers -> fully coalesced global stores, with the
// tensor-core work (2 MMA instructions per tile) and the TMA loads hidden under it.
//
// Mapping (one 128-column tile of the plane per CTA, 2-3 CTAs resident per SM overlap each other):
//   MMA M (128 TMEM lanes) = 128 consecutive plane columns n      A = proto tile, MN-major (n contiguous)
//   MMA N (TMEM columns)   = the queries, padded to 16            B = coeff, K-major (k contiguous)
//   MMA K                  = K padded to 16 (TMA zero-fills the padding)
// so that in the epilogue lane i of a warp owns plane column n0+i and a warp-wide store of one query row
// is one contiguous 128-byte (fp32) segment of `out`.
//
// Shared-memory operand layouts are the canonical UMMA 128B-swizzle layouts, produced directly by TMA
// with CU_TENSOR_MAP_SWIZZLE_128B:
//   A (MN-major): per 64-column half a box of KP rows x 128 B; 8-row groups 1024 B apart (SBO), the two
//                 halves KP*128 B apart (LBO); a K step of 16 advances the start address by 2048 B.
//   B (K-major) : rows (queries) of 128 B = 64 k-slots of which KP are used; 8-row groups 1024 B apart
//                 (SBO); a K step of 16 advances the start address by 32 B inside the swizzle atom.
#pragma once

#include <cuda.h>

#include "msda_common.cuh"

namespace msda {

constexpr int kTcTileN = 128;        // plane columns per CTA (= MMA M)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// t / n for 0 <= t < 2^23, n > 0, inv_n = 1.f / n: a multiply and two corrections instead of the ~40-instruction integer
// division sequence.  The persistent kernels decode one work item per ~1500 cycles in every role, and the three dependent
// divisions of `decode` sat on the epilogue warps' critical path (clock stamps: ~450 idle cycles between two items,
// profiles/r02_mask_kernels.md).
__device__ __forceinline__ int fast_div_small(int t, int n, float inv_n) {
  int q = __float2int_rz(static_cast<float>(t) * inv_n);
  q -= (q * n > t) ? 1 : 0;
  q += ((q + 1) * n <= t) ? 1 : 0;
  return q;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// UMMA shared-memory descriptor, 128B swizzle (layout type 2), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);                 // start address, bits [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;            // leading byte offset, bits [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;            // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                                     // version = 1
  d |= static_cast<uint64_t>(2) << 61;                                     // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = fp32, A and B both bf16 (format 1) or both fp16 (format 0), A MN-major, B K-major,
// M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_16bit_m128(uint32_t n, bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | (1u << 15) | (0u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: several can be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tma_store_3d_nocommit(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// -------------------------------------------------------------------------------------------------
// Persistent, warp-specialised version: one CTA per SM walks work items (b, 128-column tile, query chunk);
// a 4-stage TMA ring feeds the single-thread MMA issuer, two TMEM accumulators (columns 0 and 256) let the
// tensor core fill one while eight epilogue warps drain the other, so the output stream to HBM never stops.
//   warp 0      TMA producer (one lane)
//   warp 1      TMEM allocation + MMA issue (one lane)
//   warps 2..9  epilogue: warp w drains TMEM lanes 32*(w%4)..+31 (the hardware restricts a warp to the lane
//               quarter given by its index mod 4); the two warps of a quarter take alternate 32-query blocks
constexpr int kTc2Stages = 4;
constexpr int kTc2Threads = 320;
constexpr int kTc2EpiWarps = 8;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// 32 queries x 32 plane columns of one warp -> global.  Lane i owns plane column `col`.
__device__ __forceinline__ void mask_store_block(float* out_col, int64_t Ncols, int qn, bool col_ok, int lane, const float (&v)[32]) {
  if (!col_ok) return;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < qn) __stcs(out_col + static_cast<int64_t>(j) * Ncols, v[j]);
}
// bf16: neighbouring lanes trade one value per row pair so that every store is a packed bf16x2 (4 bytes per
// lane): even lanes write the even rows (columns i, i+1), odd lanes the odd rows (columns i-1, i).
__device__ __forceinline__ void mask_store_block(__nv_bfloat16* out_col, int64_t Ncols, int qn, bool col_ok, int lane, const float (&v)[32]) {
  const bool odd = lane & 1;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float give = odd ? v[j] : v[j + 1];
    const float got = __shfl_xor_sync(0xffffffffu, give, 1);
    const int row = j + (odd ? 1 : 0);
    const __nv_bfloat162 packed = odd ? __floats2bfloat162_rn(got, v[j + 1]) : __floats2bfloat162_rn(v[j], got);
    // Ncols is even and tiles start at multiples of 128, so the partner column is valid whenever this one is
    if (col_ok && row < qn)
      *reinterpret_cast<__nv_bfloat162*>(out_col + static_cast<int64_t>(row) * Ncols - (odd ? 1 : 0)) = packed;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void st_stage(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_stage(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st_stage(__half* p, float v) { *p = __float2half_rn(v); }

// Work item = (b, 128-column tile, query chunk qc): rows [qc*QS, qc*QS + rows) with QS a multiple of 32 so that the
// 32-row output boxes of one item never reach into the next item's rows; the last chunk takes the remainder and
// TMA clips its final box at Q.  QN = MMA N = rows of the largest chunk rounded up to 16.
//
// Epilogue: the first kernel stored straight from registers (one 128-byte row segment per warp instruction) and
// was limited by the number of stores eight warps can keep in flight (~9 B/clk/SM, clock stamps in
// profiles/r01e_mask_tc2_timeline.txt).  Here a chunk of 32 query rows x 128 columns is transposed through shared
// memory and leaves as ONE bulk tensor store, so a few instructions keep tens of KB in flight.
template <typename OT>
__global__ void __launch_bounds__(kTc2Threads, 1)
mask_fwd_tc2_kernel(const __grid_constant__ CUtensorMap map_proto, const __grid_constant__ CUtensorMap map_coeff,
                    const __grid_constant__ CUtensorMap map_out, int Q, int KP, int QS, int QN, int n_qchunks,
                    int n_tiles_n, int n_items, int in_fp16, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  const uint32_t a_half = static_cast<uint32_t>(KP) * 128u;
  const uint32_t b_bytes = static_cast<uint32_t>(QN) * 128u;
  const uint32_t stage_bytes = (2 * a_half + b_bytes + 1023u) & ~1023u;
  constexpr uint32_t kOutBuf = 32u * kTcTileN * sizeof(OT);             // one 32-row x 128-column output box
  uint8_t* out_stage = smem + kTc2Stages * stage_bytes;                  // [half][2 buffers]
  __shared__ __align__(8) uint64_t bars[2 * kTc2Stages + 4];
  __shared__ uint32_t s_tmem_base;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kTc2Stages + s); };
  auto bar_tfull = [&](int a) { return bar0 + 8u * (2 * kTc2Stages + a); };
  auto bar_tempty = [&](int a) { return bar0 + 8u * (2 * kTc2Stages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTc2Stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), kTc2EpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_proto) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_coeff) : "memory");
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

  // items are ordered chunk-major (all tiles of chunk 0, then chunk 1, ...): neighbouring CTAs share `coeff` rows
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
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
        const int s = i % kTc2Stages;
        const uint32_t ph = (i / kTc2Stages) & 1;
        int b, tile, qc;
        decode(item, b, tile, qc);
        mbar_wait(bar_empty(s), ph ^ 1);
        const uint32_t dst = smem_u32(smem) + s * stage_bytes;
        mbar_expect_tx(bar_full(s), 2 * a_half + b_bytes);
        tma_load_3d(dst, &map_proto, bar_full(s), tile * kTcTileN, 0, b);
        tma_load_3d(dst + a_half, &map_proto, bar_full(s), tile * kTcTileN + 64, 0, b);
        tma_load_3d(dst + 2 * a_half, &map_coeff, bar_full(s), 0, qc * QS, b);
        if (dbg && blockIdx.x == 0 && i < 16) dbg[0 * 16 + i] = clock64();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16bit_m128(static_cast<uint32_t>(QN), in_fp16 != 0);
      int i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
        const int s = i % kTc2Stages, a = i & 1;
        const uint32_t ph = (i / kTc2Stages) & 1, aph = (i >> 1) & 1;
        mbar_wait(bar_tempty(a), aph ^ 1);                     // the epilogue has drained this accumulator
        if (dbg && blockIdx.x == 0 && i < 16) dbg[1 * 16 + i] = clock64();
        mbar_wait(bar_full(s), ph);                            // operands have landed
        if (dbg && blockIdx.x == 0 && i < 16) dbg[2 * 16 + i] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(smem) + s * stage_bytes, b_addr = a_addr + 2 * a_half;
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint64_t a_desc = umma_desc_sw128(a_addr + ks * 2048u, a_half, 1024u);
          const uint64_t b_desc = umma_desc_sw128(b_addr + ks * 32u, 16u, 1024u);
          umma_bf16(tmem_base + a * 256u, a_desc, b_desc, idesc, ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty(s)) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull(a)) : "memory");
      }
    }
  } else {
    // epilogue: two independent groups ("halves") of four warps; a group owns alternate 32-row chunks of an item
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const bool is_issuer = ((warp - 2) & 3) == 0 && lane == 0;    // first warp of each group
    OT* my_stage = reinterpret_cast<OT*>(out_stage + (half * (sizeof(OT) == 2 ? 4 : 2)) * kOutBuf);
    uint32_t use = 0;                                               // passes this group has staged so far
    int i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
      const int a = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      int b, tile, qc;
      decode(item, b, tile, qc);
      const int q_begin = qc * QS;
      const int rows = (qc == n_qchunks - 1) ? (Q - q_begin) : QS;
      mbar_wait(bar_tfull(a), aph);
      if (dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 16) dbg[3 * 16 + i] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = tmem_base + a * 256u + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (sizeof(OT) == 2) {
        // 16-bit outputs: a group stages TWO 32-row chunks per pass (buffers of 8 KB: four per group, two passes in flight), so an
        // item of <= 128 rows costs each group one pair of barriers, one TMEM wait and one proxy fence instead of two of each
        // (clock stamps before / after: profiles/r02_mask_kernels.md)
        for (int q0 = half * 32; q0 < rows; q0 += 128, ++use) {
          OT* buf0 = my_stage + (use & 1) * (2 * 32 * kTcTileN);
          OT* buf1 = buf0 + 32 * kTcTileN;
          const bool two = q0 + 64 < rows;
          if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the pass before the previous one has been read
          named_bar_sync(1 + half, 128);
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(lane_base + static_cast<uint32_t>(q0), r0);
          if (two) tmem_ld32_nowait(lane_base + static_cast<uint32_t>(q0 + 64), r1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) st_stage(buf0 + j * kTcTileN + quarter * 32 + lane, __uint_as_float(r0[j]));
          if (two) {
#pragma unroll
            for (int j = 0; j < 32; ++j) st_stage(buf1 + j * kTcTileN + quarter * 32 + lane, __uint_as_float(r1[j]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar_sync(1 + half, 128);
          if (is_issuer) {
            tma_store_3d_nocommit(&map_out, smem_u32(buf0), tile * kTcTileN, q_begin + q0, b);
            if (two) tma_store_3d_nocommit(&map_out, smem_u32(buf1), tile * kTcTileN, q_begin + q0 + 64, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
      for (int q0 = half * 32; q0 < rows; q0 += 64, ++use) {
        OT* buf = my_stage + (use & 1) * (32 * kTcTileN);
        // the bulk store that last read this buffer (two chunks ago) must have drained it
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
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 16) dbg[4 * 16 + i] = clock64();
      if (lane == 0) mbar_arrive(bar_tempty(a));
    }
    if (is_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// -------------------------------------------------------------------------------------------------
// fp32 inputs on the tensor cores with fp32 accuracy ("3xTF32"): the reference trains in fp32 (AMP disabled,
// configs/R50_coco.yaml:41-42) and a single TF32 or bf16 pass would miss the 1e-4 parity bar, so every operand is split
// into  a = hi + lo  (hi = a truncated to TF32, which is what the MMA does to a raw fp32 operand; lo = a - hi, exact in fp32)
// and the product is accumulated as  hi*hi + hi*lo + lo*hi  in the fp32 TMEM accumulator (the dropped lo*lo term is 2^-22
// relative).  The kernels are in mask_tc4.cuh (forward / grad_proto), mask_tc_bwd.cuh (grad_coeff) and gemm3x.cuh (Linear).
// kind::tf32, D = fp32, A and B both K-major (32-bit operands are staged K-major: the split warps transpose proto)
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m128(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

inline size_t mask_tc2_smem_bytes(int KP, int QN, size_t out_elem) {
  const size_t stage = (2 * static_cast<size_t>(KP) * 128 + static_cast<size_t>(QN) * 128 + 1023) & ~size_t(1023);
  return 1024 + kTc2Stages * stage + (out_elem == 2 ? 8 : 4) * 32 * kTcTileN * out_elem;      // 16-bit outputs: two chunks per pass
}

}  // namespace msda
