// fp32 GEMM on the 5th-generation tensor cores as 3xTF32 (sm_100a): C[M x N] (+)= A[M x K] * B[K x N].
//
// Used for the Linear layers around the sampler (/root/reference/mdqe/models/ops/modules/ms_deform_attn.py:136-138
// value_proj + masked_fill, :143-146 sampling_offsets / sampling_grid_offsets, :157 attention_weights, :171 output_proj), which the
// reference runs as fp32 cuBLAS SGEMMs (TF32 disabled).  hi = the fp32 operand itself (the MMA truncates to TF32), lo = a - trunc(a)
// produced on chip; hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM: ~1e-6 of the exact product.
//
// Both operands may be K-major (reduction index contiguous: TMA 128B swizzle) or MN-major (M / N index contiguous: "128B swizzle,
// 32B atoms"), so the three GEMMs of a Linear layer need no transposed copies:
//   y  = x  W^T      A = x  [R][in]   K-major      B = W  [out][in]  K-major
//   dx = dy W        A = dy [R][out]  K-major      B = W  [out][in]  = [K][N] -> MN-major
//   dW = dy^T x      A = dy [R][out]  = [K][M] -> MN-major          B = x [R][in] = [K][N] -> MN-major, reduction over the rows
//                    split across CTAs, partial tiles added with TMA reduce-add stores into a zeroed dW; the column sums of dy
//                    (the bias gradient) are taken by the warps that stage dy.
// One persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2-9 epilogue (TMEM -> registers ->
// bias / row mask -> swizzled staging tile -> TMA store), warps 10-17 stage the operands.  Tile 128 x 128, reduction in chunks of
// 32, 4-stage ring, two accumulators in TMEM so the epilogue of one segment overlaps the MMAs of the next.
//
// A operand through tensor memory.  An MMA of shape M128 x N128 x K8 that takes both operands from shared memory fetches 4 KB of A
// and 4 KB of B at 128 B/clk, and producing the lo tiles costs another read + write of every operand byte: 192 KB of shared-memory
// traffic per 32-wide chunk.  Here the staging warps read the raw A tile once and write `hi` and `lo` straight into TMEM
// (tcgen05.st, lane = row of the tile, column = k; an MN-major tile is transposed by the read), the MMAs take A from there
// ([a_tmem] operand form) and only B from shared memory: 128 KB per chunk, stages of 48 KB instead of 64.  TMEM: columns 0-255
// accumulators, 256 + 64 s + {0, 32} the hi / lo chunk of stage s.  Measured against the all-shared form (profiles/r02aj_*):
// 24.3 -> 21.4 us forward, 22.5 -> 20.5 dgrad at 20 400 x 256 -> 256; the tensor pipe is busy 103 clocks per MMA in both.
//
// Work distribution (mode 2, "stream-K").  20 400 rows x 256 columns are 320 tiles on 148 SMs: 24 CTAs would compute three tiles
// while 124 compute two, and the 14 tiles of a 784-row decoder GEMM would leave 134 SMs idle for the latency of eight chunks.
// Instead the tiles x chunks "units" are dealt out as one contiguous range per CTA.  A tile whose chunks fall into several ranges
// is combined in the output itself: the CTA that holds the LAST chunks of the tile -- the first thing it does -- stores its
// partial tile (plus bias) and publishes a per-tile flag; the others wait for the flag and add theirs with TMA reduce-add stores.
// Nobody waits for work that is scheduled later than its own, so the wait cannot deadlock; the last arrival resets the flag.
#pragma once

#include "mask_tc4.cuh"

namespace msda {

constexpr int kG3Tile = 128;
constexpr int kG3Stages = 4;
constexpr int kG3SplitWarps = 8;
constexpr int kG3Threads = (2 + 8 + kG3SplitWarps) * 32;               // 576
constexpr uint32_t kG3OpBytes = kG3Tile * 128u;                        // one operand tile: 128 rows/columns x 32 fp32
constexpr uint32_t kG3StageBytes = 3 * kG3OpBytes;                     // [A raw][B hi][B lo]
constexpr uint32_t kG3TmemACol = 256;                                  // first TMEM column of the A chunks
constexpr uint32_t kG3OutBytes = kG3Tile * 128u;                       // staging: 128 rows x 32 columns
constexpr int kG3FlagSpins = 1 << 22;                                  // x 64 ns: a lost flag costs a quarter second, not a hang

enum : int { kG3ModeStore = 0, kG3ModeReduce = 1, kG3ModeStreamK = 2 };

// how often a stream-K stretch gave up waiting for its tile's flag (its reduce-add then lands on whatever the output holds: a wrong
// tile, never a hang).  Must stay 0; msda_gemm_flag_timeouts() reads it, tests/test_linear_gpu.py asserts it after the stress runs.
__device__ unsigned int g_g3_flag_timeouts = 0;

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// One stretch of work of a CTA: chunks [c0, c1) of the reduction for output tile (tm, tn).
//   out: 0 = store (+ bias), 1 = reduce-add, 2 = store (+ bias) and publish flags[tile], 3 = wait for flags[tile], then reduce-add
struct G3Seg { int tm, tn, c0, c1, out, tile; bool with_bias; };

// The same walk is made by all four roles of a CTA.  mode 0 / 1: items (split, tm, tn) dealt round-robin, the reduction cut into
// `chunks_per_split` ranges; mode 2: one contiguous range of tile-major (tile, chunk) units per CTA.
struct G3Walk {
  int mode, n_kchunks, cps, tiles_m, tiles_n, cur, end, stride;
  __device__ G3Walk(int mode_, int n_kchunks_, int cps_, int tiles_m_, int tiles_n_, int n_items)
      : mode(mode_), n_kchunks(n_kchunks_), cps(cps_), tiles_m(tiles_m_), tiles_n(tiles_n_) {
    if (mode == kG3ModeStreamK) {
      const int q = n_items / static_cast<int>(gridDim.x), r = n_items % static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
      cur = c * q + min(c, r);
      end = cur + q + (c < r ? 1 : 0);
      stride = 0;
    } else {
      cur = static_cast<int>(blockIdx.x);
      end = n_items;
      stride = static_cast<int>(gridDim.x);
    }
  }
  __device__ bool next(G3Seg& s) {
    if (cur >= end) return false;
    if (mode == kG3ModeStreamK) {
      s.tile = cur / n_kchunks;
      s.c0 = cur - s.tile * n_kchunks;
      s.c1 = min(n_kchunks, s.c0 + (end - cur));
      cur += s.c1 - s.c0;
      s.tm = s.tile / tiles_n;
      s.tn = s.tile - s.tm * tiles_n;
      s.out = s.c1 == n_kchunks ? (s.c0 == 0 ? 0 : 2) : 3;
      s.with_bias = s.out != 3;
    } else {
      const int r = cur / tiles_n, split = r / tiles_m;
      s.tn = cur - r * tiles_n;
      s.tm = r - split * tiles_m;
      s.c0 = split * cps;
      s.c1 = min(n_kchunks, s.c0 + cps);
      s.out = mode;
      s.with_bias = split == 0;
      s.tile = 0;
      cur += stride;
    }
    return true;
  }
};

// how many CTAs hold chunks of `tile` (mode 2)
__device__ __forceinline__ int g3_tile_parts(int tile, int n_kchunks, int n_units) {
  const int q = n_units / static_cast<int>(gridDim.x), r = n_units % static_cast<int>(gridDim.x);
  auto cta_of = [&](int u) { return u < r * (q + 1) ? u / (q + 1) : r + (u - r * (q + 1)) / q; };
  return cta_of(tile * n_kchunks + n_kchunks - 1) - cta_of(tile * n_kchunks) + 1;
}

// Epilogue that writes C as the sampler's paired-corner bf16 value layout (PackedLevel, msda_common.cuh) instead of a row-major fp32
// matrix: C rows are the (n, s) pixels of value_proj's output, every 32 columns one head.  packed == nullptr: the normal epilogue.
struct G3Packed {
  uint4* packed;
  const int64_t* shapes;
  const int64_t* level_start;
  int L, S, heads;
};

template <bool kAMn, bool kBMn>
__global__ void __launch_bounds__(kG3Threads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_c, const float* __restrict__ bias,
              const unsigned char* __restrict__ row_mask, float* __restrict__ col_sum_a, int* __restrict__ flags, int M, int N,
              int n_kchunks, int chunks_per_split, int tiles_m, int tiles_n, int n_items, int mode, G3Packed pk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* out_stage = smem + kG3Stages * kG3StageBytes;                // 2 x 16 KB (one per column half)
  __shared__ __align__(8) uint64_t bars[3 * kG3Stages + 4];
  __shared__ uint32_t s_tmem_base;
  __shared__ PackedLevel s_plv[kMaxLevels];                             // pk.packed only
  __shared__ int s_ptotal;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (kG3Stages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * kG3Stages + s); };
  auto bar_tfull = [&](int a) { return bar0 + 8u * (3 * kG3Stages + a); };
  auto bar_tempty = [&](int a) { return bar0 + 8u * (3 * kG3Stages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kG3Stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), kG3SplitWarps); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();            // launched with programmatic stream serialization: everything above overlaps the previous kernel's drain
  if (pk.packed != nullptr) {                                            // (uniform) level table of the packed layout for the epilogue
    stage_packed_levels(s_plv, &s_ptotal, pk.shapes, pk.level_start, pk.L, pk.S);
    __syncthreads();
  }

  G3Walk walk(mode, n_kchunks, chunks_per_split, tiles_m, tiles_n, n_items);
  G3Seg sg;

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      while (walk.next(sg)) {
        for (int kc = sg.c0; kc < sg.c1; ++kc, ++i) {
          const int s = i % kG3Stages;
          const uint32_t ph = (i / kG3Stages) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t dst = smem_u32(smem) + s * kG3StageBytes;
          mbar_expect_tx(bar_full(s), 2 * kG3OpBytes);
          if constexpr (kAMn) {
            for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * 4096u, &map_a, bar_full(s), sg.tm * kG3Tile + j * 32, kc * 32, 0);
          } else {
            tma_load_3d(dst, &map_a, bar_full(s), kc * 32, sg.tm * kG3Tile, 0);
          }
          if constexpr (kBMn) {
            for (int j = 0; j < 4; ++j) tma_load_3d(dst + kG3OpBytes + j * 4096u, &map_b, bar_full(s), sg.tn * kG3Tile + j * 32, kc * 32, 0);
          } else {
            tma_load_3d(dst + kG3OpBytes, &map_b, bar_full(s), kc * 32, sg.tn * kG3Tile, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kG3Tile, 0u, kBMn ? 1u : 0u);
      int i = 0, it = 0;
      for (; walk.next(sg); ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(bar_tempty(a), aph ^ 1);
        uint32_t acc = 0;
        for (int kc = sg.c0; kc < sg.c1; ++kc, ++i) {
          const int s = i % kG3Stages;
          const uint32_t ph = (i / kG3Stages) & 1;
          mbar_wait(bar_ready(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_hi = smem_u32(smem) + s * kG3StageBytes + kG3OpBytes, b_lo = b_hi + kG3OpBytes;
          const uint32_t a_hi = tmem_base + kG3TmemACol + static_cast<uint32_t>(s) * 64u, a_lo = a_hi + 32u;
          const uint32_t a_sel[3] = {a_hi, a_hi, a_lo}, b_sel[3] = {b_hi, b_lo, b_hi};     // hi*hi + hi*lo + lo*hi
          for (int term = 0; term < 3; ++term)
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t b_desc = kBMn ? umma_desc(b_sel[term] + ks * 1024u, 4096u, 512u, 1u) : umma_desc(b_sel[term] + ks * 32u, 16u, 1024u, 2u);
              umma_tf32_ta(tmem_base + a * 128u, a_sel[term] + ks * 8u, b_desc, idesc, acc);
              acc = 1;
            }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty(s)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull(a)) : "memory");
      }
    }
  } else if (warp >= 10) {
    // ---- operand staging: the A chunk goes to TMEM as hi / lo columns (lane = row of the tile), B gets its lo twin in shared memory
    const uint32_t t = threadIdx.x - 10 * 32;                              // 0 .. 255
    const int quarter = warp & 3, khalf = (warp - 10) >> 2;                 // TMEM lane quarter of this warp; which 16 of the chunk's 32 k
    const int row = quarter * 32 + lane;
    int i = 0;
    while (walk.next(sg)) {
      float col_sum = 0.f;                                                  // kAMn: sum over this segment's reduction range of A[., row]
      for (int kc = sg.c0; kc < sg.c1; ++kc, ++i) {
        const int s = i % kG3Stages;
        const uint32_t ph = (i / kG3Stages) & 1;
        mbar_wait(bar_full(s), ph);
        uint8_t* st = smem + s * kG3StageBytes;
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
        const uint4* b_hi = reinterpret_cast<const uint4*>(st + kG3OpBytes);
        uint4* b_lo = reinterpret_cast<uint4*>(st + 2 * kG3OpBytes);
#pragma unroll
        for (uint32_t k = t; k < kG3OpBytes / 16; k += kG3SplitWarps * 32) {
          const uint4 v = b_hi[k];
          b_lo[k] = make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(s));
      }
      if constexpr (kAMn) {
        // column sums of A over this segment: the bias gradient of a Linear layer when A = grad_y (tn == 0 segments only)
        if (col_sum_a != nullptr && sg.tn == 0 && sg.tm * kG3Tile + row < M) atomicAdd(col_sum_a + sg.tm * kG3Tile + row, col_sum);
      }
    }
  } else {
    // ---- epilogue: warp (quarter, half) drains TMEM lanes [32*quarter, +32) x columns [64*half, +64) in two 32-column groups
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const bool is_issuer = quarter == ((2 + 4 * half) & 3) && lane == 0;     // first warp of each half
    uint8_t* my_stage = out_stage + half * kG3OutBytes;
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (; walk.next(sg); ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m = sg.tm * kG3Tile + row_in_tile;
      const bool masked = row_mask != nullptr && m < M && row_mask[m] != 0;
      if (bias != nullptr && sg.with_bias && warp == 2 && sg.tn * kG3Tile + lane * 4 < N)      // the tile's 128 bias values: in L1 before the accumulator is
        asm volatile("prefetch.global.L1 [%0];" ::"l"(bias + sg.tn * kG3Tile + lane * 4));    // (warp 2 covers 512 bytes)
      mbar_wait(bar_tfull(a), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = tmem_base + a * 128u + (static_cast<uint32_t>(quarter * 32) << 16);
      if (pk.packed != nullptr) {
        // ---- packed value epilogue: this thread's row is pixel (n, s); each 32-column group is one head's 32 channels = 64 bytes of
        // bf16, which is the RIGHT half of line xp = x and the LEFT half of line xp = x + 1 of its level row (zeros beyond the level)
        int64_t line = -1;                                                  // line of (y, xp = x), head 0; -1: row outside the table
        bool first_x = false, last_x = false;
        if (m < M) {
          const int n = m / pk.S, s_ = m - n * pk.S;
          for (int l = 0; l < pk.L; ++l) {
            const PackedLevel lv = s_plv[l];
            const int q = s_ - lv.start;
            if (q >= 0 && q < lv.H * lv.W) {
              const int y = q / lv.W, x = q - y * lv.W;
              line = (static_cast<int64_t>(n) * 2 * pk.S + lv.pstart + y * (lv.W + 1) + x) * pk.heads;
              first_x = x == 0;
              last_x = x == lv.W - 1;
              break;
            }
          }
        }
        // The 64-byte rows go through the staging tile (16-byte pieces XOR-swizzled: conflict-free both ways) so that four
        // consecutive lanes write one contiguous 64-byte half line -- full sectors -- instead of every lane its own 16 bytes 1 KB apart.
        uint2* s_line = reinterpret_cast<uint2*>(my_stage + 8192);           // per row of the tile: {line of head 0 (low 32 bits), flags}
        const int t128 = (warp - 2 - 4 * half) * 32 + lane;                  // 0 .. 127 inside this column half
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int col0 = half * 64 + g * 32, n0 = sg.tn * kG3Tile + col0;
          float v[32];
          tmem_ld32(lane_base + static_cast<uint32_t>(col0), v);
          named_bar_sync(1 + half, 128);                                    // the previous group's copies have left the staging tile
          {
            uint8_t* row = my_stage + row_in_tile * 64;
            const int f = (row_in_tile >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t u[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = 8 * j + 2 * e;
                const bool in_n = n0 + c < N;
                const float lo_ = (masked || !in_n) ? 0.f : v[c] + (bias != nullptr ? __ldg(bias + n0 + c) : 0.f);
                const float hi_ = (masked || !in_n) ? 0.f : v[c + 1] + (bias != nullptr ? __ldg(bias + n0 + c + 1) : 0.f);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo_, hi_);
                u[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(row + ((j ^ f) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
            }
            if (g == 0) s_line[row_in_tile] = make_uint2(static_cast<uint32_t>(line), (line < 0 ? 4u : 0u) | (first_x ? 1u : 0u) | (last_x ? 2u : 0u));
          }
          named_bar_sync(1 + half, 128);
          if (n0 < N) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = t128 + 128 * i, r = idx >> 2, pc = idx & 3;
              const uint2 li = s_line[r];
              if (li.y & 4u) continue;
              const uint4 q4 = *reinterpret_cast<const uint4*>(my_stage + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
              uint4* right = pk.packed + (static_cast<int64_t>(li.x) + (n0 >> 5)) * 8 + 4 + pc;      // line xp = x: pieces 4..7
              uint4* left = right + static_cast<int64_t>(pk.heads) * 8 - 4;                          // line xp = x + 1: pieces 0..3
              *right = q4;
              *left = q4;
              if (li.y & 1u) right[-4] = z;                                 // left half of line xp = 0
              if (li.y & 2u) left[4] = z;                                   // right half of line xp = W
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(a));
        continue;
      }
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int col0 = half * 64 + g * 32;                                // column of the tile
        const int n0 = sg.tn * kG3Tile + col0;
        if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        named_bar_sync(1 + half, 128);
        float v[32];
        tmem_ld32(lane_base + static_cast<uint32_t>(col0), v);
        if (bias != nullptr && sg.with_bias) {
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
        if (is_issuer) {
          if (sg.out == 3 && g == 0) {                                      // the tile's last chunks (another CTA's first work) must be in place
            const volatile int* f = flags + sg.tile;
            int spins = 0;
            for (; *f < 2 && spins < kG3FlagSpins; ++spins) __nanosleep(64);
            if (spins >= kG3FlagSpins) atomicAdd(&g_g3_flag_timeouts, 1u);
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (n0 < N) {
            if (sg.out & 1) tma_reduce_add_3d(&map_c, smem_u32(my_stage), n0, sg.tm * kG3Tile, 0);
            else tma_store_3d(&map_c, smem_u32(my_stage), n0, sg.tm * kG3Tile, 0);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // the accumulator is free again: before the flag traffic, which
      __syncwarp();                                                         // waits for the bulk stores to land
      if (lane == 0) mbar_arrive(bar_tempty(a));
      if (is_issuer && sg.out >= 2) {
        if (sg.out == 2) {                                                  // publish: both column halves stored -> flag = 2
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          __threadfence();
          atomicAdd(flags + sg.tile, 1);
        } else if (atomicAdd(flags + sg.tile, 1) == 2 * g3_tile_parts(sg.tile, n_kchunks, n_items) - 1) {
          atomicExch(flags + sg.tile, 0);                                   // every part has passed the wait: leave the flag clean for the next launch
        }
      }
    }
    if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tiles have been read; the writes complete with the grid
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

constexpr size_t kG3SmemBytes = 1024 + kG3Stages * kG3StageBytes + 2 * kG3OutBytes;

}  // namespace msda
