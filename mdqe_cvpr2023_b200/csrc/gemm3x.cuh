// fp32 GEMM on the 5th-generation tensor cores as 3xTF32 (sm_100a): C[M x N] (+)= A[M x K] * B[K x N].
//
// Used for the Linear layers around the sampler (/root/reference/mdqe/models/ops/modules/ms_deform_attn.py:136-138
// value_proj + masked_fill, :143-146 sampling_offsets / sampling_grid_offsets, :157 attention_weights, :171 output_proj), which the
// reference runs as fp32 cuBLAS SGEMMs (TF32 disabled).  hi = the fp32 operand itself (the MMA truncates to TF32), lo = a - trunc(a)
// produced on chip; hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM: ~1e-6 of the exact product.
//
// Both operands may be K-major (reduction index contiguous: TMA 128B swizzle, UMMA layout type 2) or MN-major (M / N index
// contiguous: "128B swizzle, 32B atoms", layout type 1), so the three GEMMs of a Linear layer need no transposed copies:
//   y  = x  W^T      A = x  [R][in]   K-major      B = W  [out][in]  K-major
//   dx = dy W        A = dy [R][out]  K-major      B = W  [out][in]  = [K][N] -> MN-major
//   dW = dy^T x      A = dy [R][out]  = [K][M] -> MN-major          B = x [R][in] = [K][N] -> MN-major, reduction over the rows
//                    split across CTAs, partial tiles added with TMA reduce-add stores into a zeroed dW.
// One persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2-9 epilogue (TMEM -> registers ->
// bias / row mask -> swizzled staging tile -> TMA store), warps 10-17 produce the lo tiles.  Tile 128 x 128, reduction in
// chunks of 32, 3-stage ring, two accumulators in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// kATm (A operand through tensor memory): an MMA of shape M128 x N128 x K8 fetches 4 KB of A and 4 KB of B from shared memory at
// 128 B/clk -- 64 clocks, as long as its math -- and the lo tiles cost another read + write of every operand byte, so the all-shared
// form is bound by shared-memory bandwidth (192 KB per 32-wide chunk = 1536 clocks against 813 clocks of tensor-core time).  With
// kATm the split warps read the raw A tile once and write `hi` and `lo` straight into TMEM (tcgen05.st, lane = row, column = k);
// the MMAs take A from there ([a_tmem] operand form) and only B from shared memory: 128 KB per chunk, and the stage shrinks from
// 64 to 48 KB (4 stages instead of 3).  TMEM: columns 0-255 accumulators, 256 + 64 s + {0, 32} the hi / lo chunk of stage s.
#pragma once

#include "mask_tc4.cuh"

namespace msda {

constexpr int kG3Tile = 128;
constexpr int kG3Stages = 3;                                           // A and B in shared memory
constexpr int kG3StagesTm = 4;                                         // A through tensor memory
constexpr int kG3SplitWarps = 8;
constexpr int kG3Threads = (2 + 8 + kG3SplitWarps) * 32;               // 576
constexpr uint32_t kG3OpBytes = kG3Tile * 128u;                        // one operand tile: 128 rows/columns x 32 fp32
constexpr uint32_t kG3StageBytes = 4 * kG3OpBytes;                     // [A hi][A lo][B hi][B lo]
constexpr uint32_t kG3StageBytesTm = 3 * kG3OpBytes;                   // [A raw][B hi][B lo]
constexpr uint32_t kG3TmemACol = 256;                                  // first TMEM column of the A chunks (kATm)
constexpr uint32_t kG3OutBytes = kG3Tile * 128u;                       // staging: 128 rows x 32 columns

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

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

template <bool kAMn, bool kBMn, bool kATm>
__global__ void __launch_bounds__(kG3Threads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias,
              const unsigned char* __restrict__ row_mask, float* __restrict__ col_sum_a, int M, int N, int n_kchunks, int chunks_per_split,
              int tiles_m, int tiles_n, int n_items, int reduce) {
  constexpr int kStages = kATm ? kG3StagesTm : kG3Stages;
  constexpr uint32_t kStageBytes = kATm ? kG3StageBytesTm : kG3StageBytes;
  constexpr uint32_t kBOff = kATm ? kG3OpBytes : 2 * kG3OpBytes;        // B hi inside a stage; B lo follows it
  constexpr uint32_t kTmemCols = kATm ? 512u : 256u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* out_stage = smem + kStages * kStageBytes;                    // 2 x 16 KB (one per column half)
  __shared__ __align__(8) uint64_t bars[3 * kStages + 4];
  __shared__ uint32_t s_tmem_base;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
  auto bar_tfull = [&](int a) { return bar0 + 8u * (3 * kStages + a); };
  auto bar_tempty = [&](int a) { return bar0 + 8u * (3 * kStages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), kG3SplitWarps); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  // item -> (split, row tile, column tile); column tiles of one row tile are neighbours so that A is re-read from L2
  const float inv_tiles_n = 1.f / static_cast<float>(tiles_n), inv_tiles_m = 1.f / static_cast<float>(tiles_m);
  const bool small_items = n_items < (1 << 23);
  auto decode = [&](int item, int& split, int& tm, int& tn) {
    if (small_items) {
      const int r = fast_div_small(item, tiles_n, inv_tiles_n);
      tn = item - r * tiles_n;
      split = fast_div_small(r, tiles_m, inv_tiles_m);
      tm = r - split * tiles_m;
    } else {
      tn = item % tiles_n;
      const int r = item / tiles_n;
      tm = r % tiles_m;
      split = r / tiles_m;
    }
  };
  auto chunk_range = [&](int split, int& c0, int& c1) {
    c0 = split * chunks_per_split;
    c1 = min(n_kchunks, c0 + chunks_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int split, tm, tn, c0, c1;
        decode(item, split, tm, tn);
        chunk_range(split, c0, c1);
        for (int kc = c0; kc < c1; ++kc, ++i) {
          const int s = i % kStages;
          const uint32_t ph = (i / kStages) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t dst = smem_u32(smem) + s * kStageBytes;
          mbar_expect_tx(bar_full(s), 2 * kG3OpBytes);
          if constexpr (kAMn) {
            for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * 4096u, &map_a, bar_full(s), tm * kG3Tile + j * 32, kc * 32, 0);
          } else {
            tma_load_3d(dst, &map_a, bar_full(s), kc * 32, tm * kG3Tile, 0);
          }
          if constexpr (kBMn) {
            for (int j = 0; j < 4; ++j) tma_load_3d(dst + kBOff + j * 4096u, &map_b, bar_full(s), tn * kG3Tile + j * 32, kc * 32, 0);
          } else {
            tma_load_3d(dst + kBOff, &map_b, bar_full(s), kc * 32, tn * kG3Tile, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kG3Tile, (kAMn && !kATm) ? 1u : 0u, kBMn ? 1u : 0u);
      int i = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int split, tm, tn, c0, c1;
        decode(item, split, tm, tn);
        chunk_range(split, c0, c1);
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(bar_tempty(a), aph ^ 1);
        uint32_t acc = 0;
        for (int kc = c0; kc < c1; ++kc, ++i) {
          const int s = i % kStages;
          const uint32_t ph = (i / kStages) & 1;
          mbar_wait(bar_ready(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(smem) + s * kStageBytes, a_lo = a_hi + kG3OpBytes;
          const uint32_t b_hi = a_hi + kBOff, b_lo = b_hi + kG3OpBytes;
          const uint32_t a_sel[3] = {a_hi, a_hi, a_lo}, b_sel[3] = {b_hi, b_lo, b_hi};     // hi*hi + hi*lo + lo*hi
          const uint32_t a_tm = tmem_base + kG3TmemACol + static_cast<uint32_t>(s) * 64u;   // kATm: hi at +0, lo at +32
          for (int term = 0; term < 3; ++term)
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t b_desc = kBMn ? umma_desc(b_sel[term] + ks * 1024u, 4096u, 512u, 1u) : umma_desc(b_sel[term] + ks * 32u, 16u, 1024u, 2u);
              if constexpr (kATm) {
                umma_tf32_ta(tmem_base + a * 128u, a_tm + (term == 2 ? 32u : 0u) + ks * 8u, b_desc, idesc, acc);
              } else {
                const uint64_t a_desc = kAMn ? umma_desc(a_sel[term] + ks * 1024u, 4096u, 512u, 1u) : umma_desc(a_sel[term] + ks * 32u, 16u, 1024u, 2u);
                umma_tf32(tmem_base + a * 128u, a_desc, b_desc, idesc, acc);
              }
              acc = 1;
            }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty(s)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull(a)) : "memory");
      }
    }
  } else if (warp >= 10) {
    // ---- lo tiles (element-wise, layout-agnostic); kATm: the A chunk goes to TMEM as hi / lo columns, lane = row of the tile
    const uint32_t t = threadIdx.x - 10 * 32;                              // 0 .. 255
    const int quarter = warp & 3, khalf = (warp - 10) >> 2;                 // TMEM lane quarter of this warp; which 16 of the chunk's 32 k
    const int row = quarter * 32 + lane;
    float col_sum = 0.f;                                                    // kATm && kAMn: sum over the reduction index of A[., row]
    int i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int split, tm, tn, c0, c1;
      decode(item, split, tm, tn);
      chunk_range(split, c0, c1);
      for (int kc = c0; kc < c1; ++kc, ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        mbar_wait(bar_full(s), ph);
        uint8_t* st = smem + s * kStageBytes;
        if constexpr (kATm) {
          uint32_t hi[16], lo[16];
          if constexpr (kAMn) {
            // boxes {32 m, 32 k}: row k = 128 bytes holding 32 m, 32-byte chunk index XOR (k & 3)  [128B swizzle, 32B atoms]
            const uint8_t* box = st + quarter * 4096;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = khalf * 16 + j;
              hi[j] = *reinterpret_cast<const uint32_t*>(box + k * 128 + ((((lane >> 3) ^ (k & 3)) << 5) | ((lane & 7) << 2)));
              lo[j] = tf32_lo(hi[j]);
              col_sum += __uint_as_float(hi[j]);
            }
          } else {
            // K-major rows of 32 k (128 bytes), 16-byte chunk index XOR (row & 7)  [128B swizzle]
            const uint8_t* a_row = st + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 v = *reinterpret_cast<const uint4*>(a_row + ((((khalf * 4 + j) ^ (row & 7))) << 4));
              hi[4 * j] = v.x; hi[4 * j + 1] = v.y; hi[4 * j + 2] = v.z; hi[4 * j + 3] = v.w;
              lo[4 * j] = tf32_lo(v.x); lo[4 * j + 1] = tf32_lo(v.y); lo[4 * j + 2] = tf32_lo(v.z); lo[4 * j + 3] = tf32_lo(v.w);
            }
          }
          const uint32_t a_tm = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kG3TmemACol + static_cast<uint32_t>(s) * 64u + khalf * 16u;
          tmem_st16(a_tm, hi);
          tmem_st16(a_tm + 32u, lo);
        }
#pragma unroll
        for (int op = kATm ? 1 : 0; op < 2; ++op) {
          const uint4* hi = reinterpret_cast<const uint4*>(st + (op == 0 ? 0u : kBOff));
          uint4* lo = reinterpret_cast<uint4*>(st + (op == 0 ? 0u : kBOff) + kG3OpBytes);
#pragma unroll
          for (uint32_t k = t; k < kG3OpBytes / 16; k += kG3SplitWarps * 32) {
            const uint4 v = hi[k];
            uint4 l;
            l.x = tf32_lo(v.x);
            l.y = tf32_lo(v.y);
            l.z = tf32_lo(v.z);
            l.w = tf32_lo(v.w);
            lo[k] = l;
          }
        }
        if constexpr (kATm) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if constexpr (kATm) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(s));
      }
      if constexpr (kATm && kAMn) {
        // column sums of A over this item's reduction range: the bias gradient of a Linear layer when A = grad_y (tn == 0 items only)
        if (col_sum_a != nullptr && tn == 0 && tm * kG3Tile + row < M) atomicAdd(col_sum_a + tm * kG3Tile + row, col_sum);
        col_sum = 0.f;
      }
    }
  } else {
    // ---- epilogue: warp (quarter, half) drains TMEM lanes [32*quarter, +32) x columns [64*half, +64) in two 32-column groups
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const bool is_issuer = quarter == ((2 + 4 * half) & 3) && lane == 0;     // first warp of each half
    uint8_t* my_stage = out_stage + half * kG3OutBytes;
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int split, tm, tn;
      decode(item, split, tm, tn);
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m = tm * kG3Tile + row_in_tile;
      const bool masked = row_mask != nullptr && m < M && row_mask[m] != 0;
      mbar_wait(bar_tfull(a), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = tmem_base + a * 128u + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int col0 = half * 64 + g * 32;                                // column of the tile
        const int n0 = tn * kG3Tile + col0;
        if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        named_bar_sync(1 + half, 128);
        float v[32];
        tmem_ld32(lane_base + static_cast<uint32_t>(col0), v);
        if (bias != nullptr && split == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
        }
        if (masked) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        // staging tile [128 rows][32 columns] in the 128B-swizzle layout of the output map: conflict-free 16-byte stores
        uint8_t* row = my_stage + row_in_tile * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<float4*>(row + ((c16 ^ (row_in_tile & 7)) * 16)) = make_float4(v[4 * c16], v[4 * c16 + 1], v[4 * c16 + 2], v[4 * c16 + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(1 + half, 128);
        if (is_issuer && n0 < N) {
          if (reduce) tma_reduce_add_3d(&map_c, smem_u32(my_stage), n0, tm * kG3Tile, 0);
          else tma_store_3d(&map_c, smem_u32(my_stage), n0, tm * kG3Tile, 0);
        }
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

constexpr size_t kG3SmemBytes = 1024 + kG3Stages * kG3StageBytes + 2 * kG3OutBytes;
constexpr size_t kG3SmemBytesTm = 1024 + kG3StagesTm * kG3StageBytesTm + 2 * kG3OutBytes;

}  // namespace msda
