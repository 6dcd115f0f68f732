// Second-generation fast kernels: same decomposition as msda_fast.cuh (one warp per (n,q,m) pair, one
// lane group per bilinear corner, 16-byte channel vectors) with the instruction overhead removed.
// ncu on the first generation showed the forward to be ISSUE bound (75 % issue-active, 468 warp
// instructions per pair, only ~130 of them loads/FMAs; profiles/r01a_summary.md).  Changes:
//   * L*P is a template parameter (8, 12, 16: every MDQE call), the corner loop is fully unrolled and
//     its shared-memory addresses are immediates;
//   * slot records hold the row offset in 16-byte units, so an address is ONE 32x32->64 multiply-add;
//     they are stored corner-major so one LDS.128 fetches the records of two consecutive iterations;
//   * (n, m) of a pair come from multiply-shift division by host-computed magic numbers instead of
//     two 64-bit divisions;
//   * backward dot partials go to a transposed, padded tile [lane][sample] (conflict-free, immediate
//     offsets) and phase 3 folds them per corner.
#pragma once

#include <type_traits>

#include "msda_fast.cuh"

namespace msda {


#ifdef MSDA_DBG_MASK
// timing experiments only (tools/whatif_bench.py builds a separate library with -DMSDA_DBG_MASK): per-level bits that
// drop work from the kernels -- bits 0-3 forward gathers, 4-7 backward reductions, 8-11 backward gathers + reductions
__device__ int g_dbg_mask;
#define MSDA_DBG_SKIP(shift, smp) (((dbg_mask >> ((shift) + ((smp) >> 2))) & 1) != 0)      // P = 4 only
#define MSDA_DBG_LOAD const int dbg_mask = g_dbg_mask;
#else
#define MSDA_DBG_SKIP(shift, smp) false
#define MSDA_DBG_LOAD
#endif

struct FastDiv {              // q = x / d for 0 <= x < 2^31 (host: make_fastdiv)
  uint32_t d, mul, shr;
};
__device__ __forceinline__ uint32_t fd_div(uint32_t x, const FastDiv f) {
  return f.d == 1 ? x : (__umulhi(x, f.mul) >> f.shr);
}

// Which pairs a CTA walks, and in which order.  Linear (hrun = 0): CTA b owns pairs [b * chunk, (b + 1) * chunk) of the
// flattened (n, q, m) index, i.e. chunk / M consecutive queries with all their heads.  Head-run (hrun = 1): CTA b owns ONE head
// m = b % M of `chunk` consecutive queries; its warps then sweep neighbouring queries of the same head together, so the corner
// rows they gather (same head => same 128-byte lines when the samples land in the same cells) are reused out of L1 while they
// are still there.  `begin`/`end` delimit the CTA's virtual pair indices, pair(v) maps one to the flattened pair index.
struct PairMap {
  uint32_t begin, end, base, stride, head;
  __device__ __forceinline__ PairMap(int hrun, uint32_t block, int chunk_pairs, uint32_t n_pairs, const FastDiv div_m) {
    if (hrun) {
      const uint32_t run = fd_div(block, div_m);
      head = block - run * div_m.d;
      base = run * static_cast<uint32_t>(chunk_pairs);
      stride = div_m.d;
      begin = 0;
      const uint32_t nq_total = fd_div(n_pairs, div_m);
      end = min(static_cast<uint32_t>(chunk_pairs), nq_total - base);
    } else {
      begin = block * static_cast<uint32_t>(chunk_pairs);
      end = min(n_pairs, begin + static_cast<uint32_t>(chunk_pairs));
      base = 0; stride = 1; head = 0;
    }
  }
  __device__ __forceinline__ uint32_t pair(uint32_t v) const { return (base + v) * stride + head; }
};

// Fused sampler prologue (ms_deform_attn.py:142-161 folded into the kernel): instead of ready-made
// sampling locations and softmax-normalised weights the kernel gets what the module's Linear layers produce --
// raw offsets (in the `loc` slot) and raw attention logits (in the `aw` slot) -- plus the reference points, and does
//   aw  = softmax(logits) over the L*P samples of a (query, head)
//   loc = ref_xy + offsets / scale                                                         (mode 0: pred_offsets)
//   loc = ref_xy + (grid[m,l,p] * 0.5 * ref_wh + clamp(offsets, +-ref_wh * scale)) / scale   (mode 1: box-scaled grid)
// itself; the backward returns d/d offsets and d/d logits.  ref == nullptr: plain operator.
struct FusedArgs {
  const float* ref;     // [N*Lq, R]
  const float* grid;    // [M, L, P, 2] (mode 1) or nullptr
  int R;                // 2 (cx, cy) or 4 (cx, cy, w, h)
  int mode;
  float scale;          // the module's self.scale (8)
  int row_stride;       // 0: offsets [N*Lq, M*L*P*2] and logits [N*Lq, M*L*P] are dense tensors of their own; > 0 (even): both are
                        // column ranges of one [N*Lq, row_stride] matrix -- the output of ONE Linear layer over the concatenated
                        // sampling_offsets / attention_weights weights -- and so are their gradients
};

// softmax over the LP-lane segment that holds one (query, head) pair; every lane of the warp must call this
template <int LP>
__device__ __forceinline__ float segment_softmax(float logit, bool valid) {
  float mx = valid ? logit : -INFINITY;
#pragma unroll
  for (int o = LP / 2; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = valid ? expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = LP / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return valid ? e / sum : 0.f;
}
template <int LP>
__device__ __forceinline__ float segment_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// raw offsets (x, y) -> sampling location; mask = d loc / d offset * scale (0 where the clamp is active)
__device__ __forceinline__ void fused_location(const FusedArgs& fz, uint32_t nq, uint32_t m, int ss, int LPv, float& x, float& y,
                                               float& mask_x, float& mask_y) {
  const float* rp = fz.ref + static_cast<size_t>(nq) * fz.R;
  const float inv = 1.f / fz.scale;
  mask_x = mask_y = 1.f;
  if (fz.mode == 0) {
    x = __ldg(rp) + x * inv;
    y = __ldg(rp + 1) + y * inv;
  } else {
    const float bw = __ldg(rp + 2), bh = __ldg(rp + 3);
    const float2 g = __ldg(reinterpret_cast<const float2*>(fz.grid) + (m * LPv + ss));
    const float bx = bw * fz.scale, by = bh * fz.scale;
    mask_x = (x > -bx && x < bx) ? 1.f : 0.f;
    mask_y = (y > -by && y < by) ? 1.f : 0.f;
    float cx = x > -bx ? x : -bx;  cx = cx < bx ? cx : bx;       // the two torch.where of ms_deform_attn.py:149-152
    float cy = y > -by ? y : -by;  cy = cy < by ? cy : by;
    x = __ldg(rp) + (g.x * 0.5f * bw + cx) * inv;
    y = __ldg(rp + 1) + (g.y * 0.5f * bh + cy) * inv;
  }
}

template <typename VT, int D, int LP>
struct Cfg2 : FastCfg<VT, D> {
  using B = FastCfg<VT, D>;
  static constexpr int NSG = B::NG / 4;              // sample groups (1 for fp32, 2 for bf16)
  static constexpr int SPG = LP / NSG;               // samples each lane group walks per pair
  static constexpr int QPW = 32 / LP;                // pairs prepared per warp round
  static constexpr int LPP = LP + 2;                 // corner stride in records: +16 B so that the four corner
                                                     // streams sit in different banks (one LDS.128 wavefront)
  static constexpr int NSLOT = 4 * LPP;              // slot records per pair (incl. padding)
  static constexpr int D16 = D / B::CPL;             // 16-byte units per channel row
  static constexpr bool kAllLanes = (B::G * B::NG == 32);
  static_assert(LP % NSG == 0 && SPG % 2 == 0, "sample groups must split L*P evenly");
};

// value pointer + slot offset -> address of this lane's 16-byte channel vector
template <typename VT>
__device__ __forceinline__ const VT* row_ptr(const VT* base, uint32_t off16) {
  return reinterpret_cast<const VT*>(reinterpret_cast<const char*>(base) + static_cast<uint64_t>(off16) * 16u);
}

// Phase 1 for one sample (lane): slot records in corner-major order.
//   dst[corner * LPP + s] = {off16, weight};  invalid corners: off16 = kInvalidOff, weight = 0.
template <int D16, int LPP>
__device__ __forceinline__ SampleGeom make_slots2(Slot* dst, int s, float locx, float locy, float a, const LevelInfo li,
                                                  uint32_t n, uint32_t m, int S, int M) {
  const SampleGeom g = sample_geom(locx, locy, li.H, li.W);
  const uint32_t row = static_cast<uint32_t>(M) * D16;
  const uint32_t base = (n * static_cast<uint32_t>(S) + li.start) * row + m * D16;
  const uint32_t o00 = base + static_cast<uint32_t>(g.y0 * li.W + g.x0) * row;
  const uint32_t o10 = o00 + static_cast<uint32_t>(li.W) * row;
  const float hx = 1.f - g.lx, hy = 1.f - g.ly;
  const bool v00 = g.oky0 && g.okx0, v01 = g.oky0 && g.okx1, v10 = g.oky1 && g.okx0, v11 = g.oky1 && g.okx1;
  Slot e;
  e.off = v00 ? o00 : kInvalidOff;        e.w = v00 ? hx * hy * a : 0.f;      dst[0 * LPP + s] = e;
  e.off = v01 ? o00 + row : kInvalidOff;  e.w = v01 ? g.lx * hy * a : 0.f;    dst[1 * LPP + s] = e;
  e.off = v10 ? o10 : kInvalidOff;        e.w = v10 ? hx * g.ly * a : 0.f;    dst[2 * LPP + s] = e;
  e.off = v11 ? o10 + row : kInvalidOff;  e.w = v11 ? g.lx * g.ly * a : 0.f;  dst[3 * LPP + s] = e;
  return g;
}

// The same records kept in registers (backward: they are merged across the points of a level before they are stored).
template <int D16>
__device__ __forceinline__ SampleGeom make_slot_regs(Slot (&sl)[4], float locx, float locy, float a, const LevelInfo li,
                                                     uint32_t n, uint32_t m, int S, int M) {
  const SampleGeom g = sample_geom(locx, locy, li.H, li.W);
  const uint32_t row = static_cast<uint32_t>(M) * D16;
  const uint32_t base = (n * static_cast<uint32_t>(S) + li.start) * row + m * D16;
  const uint32_t o00 = base + static_cast<uint32_t>(g.y0 * li.W + g.x0) * row;
  const uint32_t o10 = o00 + static_cast<uint32_t>(li.W) * row;
  const float hx = 1.f - g.lx, hy = 1.f - g.ly;
  const bool v00 = g.oky0 && g.okx0, v01 = g.oky0 && g.okx1, v10 = g.oky1 && g.okx0, v11 = g.oky1 && g.okx1;
  sl[0].off = v00 ? o00 : kInvalidOff;        sl[0].w = v00 ? hx * hy * a : 0.f;
  sl[1].off = v01 ? o00 + row : kInvalidOff;  sl[1].w = v01 ? g.lx * hy * a : 0.f;
  sl[2].off = v10 ? o10 : kInvalidOff;        sl[2].w = v10 ? hx * g.ly * a : 0.f;
  sl[3].off = v11 ? o10 + row : kInvalidOff;  sl[3].w = v11 ? g.lx * g.ly * a : 0.f;
  return g;
}

// Backward only: every grad_value reduction of a pair carries the same vector (the pair's grad_out row) times a scalar, so
// corners of the P points of one (pair, level) that land on the same value row can share ONE reduction with the summed
// scalar.  The P lanes of the level exchange their four (row, weight) records with xor shuffles; the lowest lane holding a
// row keeps it with the sum of all weights, the others zero theirs (a zero weight skips the reduction; the gather and the
// dot product are untouched -- skipping them as well, with the owner's dot forwarded through the dot tile, measured no
// faster).  The coarse pyramid levels, where the points of a query crowd into a few cells, lose most of their reductions
// this way -- and L2's reduction rate is what bounds the backward.  The same merge in the forward (one gather
// per distinct row) costs more in shuffles than the gathers it saves (82 -> 98 us).  All 32 lanes must call.
// Rows are matched through their cells, not their offsets: lanes exchange ONE packed cell coordinate (x0 + 8 | (y0 + 8) << 16) and
// their four weights (5 shuffles per partner instead of 8); with (ex, ey) = partner cell - own cell, own corner (dx, dy) coincides
// with partner corner (dx - ex, dy - ey) when that lies in {0,1}^2, i.e. the partner's 2x2 weights shifted by (ex, ey) are what is
// added -- two select stages instead of sixteen compares.  Cells coincide exactly when row offsets do (same pair, same level); an
// out-of-range cell is out of range for both lanes (weight 0, invalid offset), so merging it is a no-op.  `key` must be a value
// no other lane of the level can match (kNoMergeKey) for lanes without a sample and for levels wider than the 15-bit fields.
constexpr uint32_t kNoMergeKey = 0xFFFF8000u;
template <int PTS>
__device__ __forceinline__ void merge_level_slots(Slot (&sl)[4], uint32_t key, int lane) {
  const int me = lane & (PTS - 1);
  const int cx = static_cast<int>(key & 0xffffu), cy = static_cast<int>(key >> 16);
  float wsum[4];
  unsigned kill = 0;
#pragma unroll
  for (int cn = 0; cn < 4; ++cn) wsum[cn] = sl[cn].w;
#pragma unroll 1                                     // one partner at a time: 5 shuffled values live, not 15
  for (int j = 1; j < PTS; ++j) {
    const bool lower_partner = (me ^ j) < me;
    const uint32_t pk = __shfl_xor_sync(0xffffffffu, key, j);
    float pw[4];
#pragma unroll
    for (int cp = 0; cp < 4; ++cp) pw[cp] = __shfl_xor_sync(0xffffffffu, sl[cp].w, j);
    const int ex = static_cast<int>(pk & 0xffffu) - cx, ey = static_cast<int>(pk >> 16) - cy;
    const bool x0 = ex == 0, xm = ex == -1, xp = ex == 1, y0 = ey == 0, ym = ey == -1, yp = ey == 1;
    // partner row r (0: y0, 1: y0 + 1) seen from own column dx
    const float a00 = x0 ? pw[0] : (xm ? pw[1] : 0.f), a01 = x0 ? pw[1] : (xp ? pw[0] : 0.f);
    const float a10 = x0 ? pw[2] : (xm ? pw[3] : 0.f), a11 = x0 ? pw[3] : (xp ? pw[2] : 0.f);
    wsum[0] += y0 ? a00 : (ym ? a10 : 0.f);
    wsum[1] += y0 ? a01 : (ym ? a11 : 0.f);
    wsum[2] += y0 ? a10 : (yp ? a00 : 0.f);
    wsum[3] += y0 ? a11 : (yp ? a01 : 0.f);
    if (lower_partner) {                             // a lower lane holds the cell: it keeps the record
      const bool mx0 = x0 || xm, mx1 = x0 || xp, my0 = y0 || ym, my1 = y0 || yp;
      kill |= (mx0 && my0 ? 1u : 0u) | (mx1 && my0 ? 2u : 0u) | (mx0 && my1 ? 4u : 0u) | (mx1 && my1 ? 8u : 0u);
    }
  }
#pragma unroll
  for (int cn = 0; cn < 4; ++cn) sl[cn].w = ((kill >> cn) & 1u) ? 0.f : wsum[cn];
}

// ------------------------------------------------------------------------------------------ forward
#ifndef MSDA_PACKED_FMA
#define MSDA_PACKED_FMA 1       // fma.rn.f32x2 (sm_100): two channels per FMA instruction
#endif
#ifndef MSDA_FWD_LEAN_BATCH
#define MSDA_FWD_LEAN_BATCH 4
#endif
#ifndef MSDA_FWD_PAIR_FOLD
#define MSDA_FWD_PAIR_FOLD 1
#endif
// MINB = minimum resident CTAs per SM promised to ptxas: 3 leaves it 80+ registers, enough to keep a whole
// batch of gathers in flight; 6 reproduces the register-lean, load-by-load schedule.
template <typename VT, typename LT, int D, int LP, int MINB, bool GROUPED, bool FUSED = false>
__global__ void __launch_bounds__(kThreads, MINB)
msda_fwd_fast2_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                      const LT* __restrict__ aw, VT* __restrict__ out,
                      int S, int M, int L, int P, uint32_t n_pairs, int chunk_pairs, FastDiv div_m, FastDiv div_mq,
                      int G, float scale, FusedArgs fz, int hrun) {
  // G > 1: "grouped" (temporal) form -- G level tables share loc/aw, out = scale * sum_g (see msda_forward_grouped)
  // FUSED: the module's softmax / location arithmetic in phase 1 (fz) -- its own instantiation, so that the plain operator carries
  // none of it (as a run-time branch the joint-layout addressing alone cost the encoder-sized kernels 1-2 %: profiles/r02ap)
  constexpr bool kFused = FUSED && std::is_same<LT, float>::value && (LP & (LP - 1)) == 0;
  using C = Cfg2<VT, D, LP>;
  __shared__ LevelInfo s_lvl[kMaxLevels];
  __shared__ __align__(16) Slot s_slot[kWarpsPerCta][C::QPW * C::NSLOT];

  pdl_wait();
  pdl_trigger();
  stage_levels(s_lvl, shapes, level_start, G * L, S);
  __syncthreads();
  MSDA_DBG_LOAD

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / C::G, c = lane - grp * C::G;
  const bool active = C::kAllLanes || grp < C::NG;
  const int corner = grp & 3, sgrp = (grp >> 2) & (C::NSG - 1);
  Slot* my_slots = s_slot[warp];
  // this lane's slot stream: records of its corner for samples sgrp*SPG .. +SPG-1 (contiguous)
  const Slot* my_stream = my_slots + corner * C::LPP + sgrp * C::SPG;
  const VT* vlane = value + c * C::CPL;

  const PairMap pm(hrun, blockIdx.x, chunk_pairs, n_pairs, div_m);
  const uint32_t chunk_begin = pm.begin, chunk_end = pm.end;
  const int ps = (lane < C::QPW * LP) ? lane / LP : 0;        // phase-1 role: pair slot and sample of this lane
  const int ss = lane - ps * LP;
  const int lvl = ss / P;

  for (uint32_t p0 = chunk_begin + warp * C::QPW; p0 < chunk_end; p0 += kWarpsPerCta * C::QPW) {
    const int npair = static_cast<int>(min(static_cast<uint32_t>(C::QPW), chunk_end - p0));
    const bool has_sample = lane < npair * LP;
    float x = 0.f, y = 0.f, a = 0.f;
    uint32_t n = 0, m = 0, nq = 0;
    if (has_sample) {
      const uint32_t pair = pm.pair(p0 + ps);
      nq = fd_div(pair, div_m);
      m = pair - nq * div_m.d;
      n = fd_div(pair, div_mq);
      int64_t li = static_cast<int64_t>(pair) * LP + ss, ai = li;
      if constexpr (kFused) {
        if (fz.row_stride > 0) {
          const int64_t row = static_cast<int64_t>(nq) * fz.row_stride;
          li = (row >> 1) + m * LP + ss;
          ai = row + m * LP + ss;
        }
      }
      load_loc_aw2<LT>(loc, aw, li, ai, x, y, a);
    }
    if constexpr (kFused) {                         // fused prologue: (x, y) are raw offsets, a is a raw logit
      a = segment_softmax<LP>(a, has_sample);
      float mk_x, mk_y;
      if (has_sample) fused_location(fz, nq, m, ss, LP, x, y, mk_x, mk_y);
    }
    a *= scale;
    // GROUPED keeps one accumulator set per pair alive across the G level tables; the plain operator (G == 1)
    // reduces and stores each pair as soon as its corner loop ends (fewer live registers: 81 vs 94 us measured)
    // kPairFold (plain operator, two pairs per round, four corner groups of 8 lanes, 4 channels per lane): both pairs' partial
    // sums are folded across the corner groups by ONE transposing butterfly -- 4 + 2 shuffles per round instead of 2 x 8 -- which
    // leaves lane (pair, half, c) with channels 4c + 2*half + {0,1} of its pair: one 8-byte store per lane.  Every shuffle is a
    // wavefront on the L1 data pipe that binds this kernel.
    constexpr bool kPairFold = MSDA_FWD_PAIR_FOLD && !GROUPED && C::QPW == 2 && C::NG == 4 && C::G == 8 && C::CPL == 4 && C::kAllLanes;
    float acc_g[(GROUPED || kPairFold) ? C::QPW : 1][C::CPL];
    if constexpr (GROUPED || kPairFold) {
#pragma unroll
      for (int pl = 0; pl < C::QPW; ++pl)
#pragma unroll
        for (int j = 0; j < C::CPL; ++j) acc_g[pl][j] = 0.f;
    }

    // gridDim.y > 1: the level tables are split across CTAs (small clip-level calls are parallelism bound: 1568 pairs
    // are 98 CTAs); each CTA then adds its table's contribution into a zero-filled fp32 `out` with vector reductions
    const bool g_split = GROUPED && gridDim.y > 1;
    const int g_begin = g_split ? static_cast<int>(blockIdx.y) : 0;
    const int g_end = GROUPED ? (g_split ? g_begin + 1 : G) : 1;
    for (int g = g_begin; g < g_end; ++g) {
      if (has_sample) make_slots2<C::D16, C::LPP>(my_slots + ps * C::NSLOT, ss, x, y, a, s_lvl[g * L + lvl], n, m, S, M);
      __syncwarp();
#pragma unroll
      for (int pl = 0; pl < C::QPW; ++pl) {
        if (pl < npair) {
          float acc[C::CPL];
#pragma unroll
          for (int j = 0; j < C::CPL; ++j) acc[j] = GROUPED ? acc_g[GROUPED ? pl : 0][j] : 0.f;
          const uint4* stream = reinterpret_cast<const uint4*>(my_stream + pl * C::NSLOT);
          // Batches of kBatch corner rows: slot records, then the gathers, then the FMAs (with MINB = 3 ptxas keeps
          // the whole batch in flight; the default register-lean build interleaves them, which measured faster).
          constexpr int kWant = MINB >= 5 ? MSDA_FWD_LEAN_BATCH : 8;      // gathers in flight per lane: what the register budget allows
          constexpr int kBatch = (C::SPG % kWant == 0) ? kWant : ((C::SPG % 6 == 0 && kWant > 6) ? 6 : ((C::SPG % 4 == 0 && kWant > 4) ? 4 : 2));   // must divide SPG
          static_assert(C::SPG % kBatch == 0 && kBatch % 2 == 0, "batch must tile the slot stream");
#pragma unroll
          for (int b0 = 0; b0 < C::SPG; b0 += kBatch) {
            uint32_t off[kBatch];
            float w[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch / 2; ++i) {
              const uint4 two = stream[(b0 >> 1) + i];          // records of two consecutive samples
              off[2 * i] = two.x; w[2 * i] = __uint_as_float(two.y);
              off[2 * i + 1] = two.z; w[2 * i + 1] = __uint_as_float(two.w);
            }
            // a corner that is out of range has weight 0: gather and FMAs sit under ONE condition, so ptxas predicates both
            // and the destination registers need no zero-fill
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              if (active && w[u] != 0.f && !MSDA_DBG_SKIP(0, sgrp * C::SPG + b0 + u)) {
                float v[C::CPL];
                Vec16<VT>::load(row_ptr(vlane, off[u]), v);
#if MSDA_PACKED_FMA
                const float2 ww = make_float2(w[u], w[u]);
#pragma unroll
                for (int j = 0; j < C::CPL; j += 2) {
                  const float2 r = __ffma2_rn(ww, make_float2(v[j], v[j + 1]), make_float2(acc[j], acc[j + 1]));
                  acc[j] = r.x; acc[j + 1] = r.y;
                }
#else
#pragma unroll
                for (int j = 0; j < C::CPL; ++j) acc[j] = fmaf(w[u], v[j], acc[j]);
#endif
              }
            }
          }
          if constexpr (GROUPED || kPairFold) {
#pragma unroll
            for (int j = 0; j < C::CPL; ++j) acc_g[(GROUPED || kPairFold) ? pl : 0][j] = acc[j];
          } else {
#pragma unroll
            for (int k = C::NG / 2; k >= 1; k >>= 1) {
#pragma unroll
              for (int j = 0; j < C::CPL; ++j) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], k * C::G);
            }
            if (lane < C::G) Vec16<VT>::store(out + static_cast<int64_t>(pm.pair(p0 + pl)) * D + lane * C::CPL, acc);
          }
        }
      }
      __syncwarp();
    }

    if constexpr (kPairFold) {
      const bool hi_pair = (lane & 16) != 0, hi_half = (lane & 8) != 0;
      float keep[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float mine = hi_pair ? acc_g[kPairFold ? 1 : 0][j] : acc_g[0][j];
        const float send = hi_pair ? acc_g[0][j] : acc_g[kPairFold ? 1 : 0][j];
        keep[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      float two[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float mine = hi_half ? keep[2 + i] : keep[i];
        const float send = hi_half ? keep[i] : keep[2 + i];
        two[i] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      const int pl = lane >> 4;
      if (pl < npair) {
        if constexpr (std::is_same<VT, float>::value)
          *reinterpret_cast<float2*>(out + static_cast<int64_t>(pm.pair(p0 + pl)) * D + (lane & 7) * 4 + (hi_half ? 2 : 0)) = make_float2(two[0], two[1]);
      }
    }
    if constexpr (GROUPED) {
#pragma unroll
      for (int pl = 0; pl < C::QPW; ++pl) {
        if (pl < npair) {
#pragma unroll
          for (int k = C::NG / 2; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = 0; j < C::CPL; ++j) acc_g[pl][j] += __shfl_down_sync(0xffffffffu, acc_g[pl][j], k * C::G);
          }
          if (lane < C::G) {
            VT* dst = out + static_cast<int64_t>(pm.pair(p0 + pl)) * D + lane * C::CPL;
            if constexpr (std::is_same<VT, float>::value) {
              if (g_split) red_add_f32x4(dst, acc_g[pl][0], acc_g[pl][1], acc_g[pl][2], acc_g[pl][3]);
              else Vec16<VT>::store(dst, acc_g[pl]);
            } else {
              Vec16<VT>::store(dst, acc_g[pl]);
            }
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------- backward
#ifndef MSDA_BWD_BATCH
#define MSDA_BWD_BATCH 4                     // tools/bwd_variants.sh builds A/B libraries with other values
#endif
#ifndef MSDA_BWD_MINB
#define MSDA_BWD_MINB 4                      // resident CTAs per SM promised to ptxas for the plain operator: 64 registers, no spills.
#endif                                       // With the merged reductions the kernel is latency bound, not reduction bound, and
                                             // 32 resident warps beat 16 (216 -> 185 us; 1/3/5/6: 216/194/192/224, profiles/r01y)
#ifndef MSDA_BWD_MINB_GROUPED
#define MSDA_BWD_MINB_GROUPED 1              // the grouped form keeps grad_out rows of all pairs in registers (108) and its calls are small
#endif
constexpr int kBwdBatch = MSDA_BWD_BATCH;    // gathers in flight per lane (see the corner loop)

template <typename VT, typename LT, int D, int LP, bool GROUPED, bool FUSED = false>
__global__ void __launch_bounds__(kThreads, GROUPED ? MSDA_BWD_MINB_GROUPED : MSDA_BWD_MINB)
msda_bwd_fast2_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                      const LT* __restrict__ aw, const VT* __restrict__ grad_out,
                      float* __restrict__ grad_value, LT* __restrict__ grad_loc, LT* __restrict__ grad_aw,
                      int S, int M, int L, int P, uint32_t n_pairs, int chunk_pairs, FastDiv div_m, FastDiv div_mq,
                      int G, float scale, FusedArgs fz, int merge, int hrun) {
  constexpr bool kFused = FUSED && std::is_same<LT, float>::value && (LP & (LP - 1)) == 0;    // see the forward kernel
  using C = Cfg2<VT, D, LP>;
  // <grad_out, corner row> per (corner, sample): [4][40] floats per warp.  The partials are folded inside the
  // corner group with shuffles first, so the tile stays tiny and shared memory stays small: the first version
  // kept per-lane partials (34 KB per CTA), which left only ~16 KB of L1 per SM and cost the gathers their hit rate.
  // stride 40 = 8 (mod 32): the four corner groups' stores (lane c of group g writes entry g * stride + c) land in disjoint bank
  // octets, and the sample lanes' reads (entry e * stride + lane) are conflict-free for any stride; 33 made the stores 4-way
  // conflicted (ncu: 6 of 8 store wavefronts per pair excessive)
  constexpr int kDotStride = 40;
  __shared__ LevelInfo s_lvl[kMaxLevels];
  __shared__ __align__(16) Slot s_slot[kWarpsPerCta][C::QPW * C::NSLOT];
  __shared__ float s_dot[kWarpsPerCta][4 * kDotStride];

  pdl_wait();
  pdl_trigger();
  stage_levels(s_lvl, shapes, level_start, G * L, S);
  __syncthreads();
  MSDA_DBG_LOAD

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / C::G, c = lane - grp * C::G;
  const bool active = C::kAllLanes || grp < C::NG;
  const int corner = grp & 3, sgrp = (grp >> 2) & (C::NSG - 1);
  Slot* my_slots = s_slot[warp];
  const Slot* my_stream = my_slots + corner * C::LPP + sgrp * C::SPG;
  const VT* vlane = value + c * C::CPL;
  float* gvlane = grad_value + c * C::CPL;
  // dot of (corner, sample): written by lane 0 of each corner group, read by the sample lanes
  float* dot_w = s_dot[warp] + corner * kDotStride + sgrp * C::SPG;
  const float* dot_r = s_dot[warp] + lane;

  const PairMap pm(hrun, blockIdx.x, chunk_pairs, n_pairs, div_m);
  const uint32_t chunk_begin = pm.begin, chunk_end = pm.end;
  const int ps = (lane < C::QPW * LP) ? lane / LP : 0;
  const int ss = lane - ps * LP;
  const int lvl = ss / P;

  for (uint32_t p0 = chunk_begin + warp * C::QPW; p0 < chunk_end; p0 += kWarpsPerCta * C::QPW) {
    const int npair = static_cast<int>(min(static_cast<uint32_t>(C::QPW), chunk_end - p0));
    const bool has_sample = lane < npair * LP;
    float x = 0.f, y = 0.f, a = 0.f, a_raw = 0.f, mk_x = 1.f, mk_y = 1.f;
    uint32_t n = 0, m = 0, nq = 0;
    int64_t si = 0, li = 0;                          // this lane's sample: index into aw / loc (in pairs) and their gradients
    if (has_sample) {
      const uint32_t pair = pm.pair(p0 + ps);
      nq = fd_div(pair, div_m);
      m = pair - nq * div_m.d;
      n = fd_div(pair, div_mq);
      si = li = static_cast<int64_t>(pair) * LP + ss;
      if constexpr (kFused) {
        if (fz.row_stride > 0) {
          const int64_t row = static_cast<int64_t>(nq) * fz.row_stride;
          li = (row >> 1) + m * LP + ss;
          si = row + m * LP + ss;
        }
      }
      load_loc_aw2<LT>(loc, aw, li, si, x, y, a);
    }
    if constexpr (kFused) {                         // fused prologue: (x, y) are raw offsets, a is a raw logit
      a = segment_softmax<LP>(a, has_sample);
      if (has_sample) fused_location(fz, nq, m, ss, LP, x, y, mk_x, mk_y);
    }
    a_raw = a;
    a *= scale;                                     // d out / d value carries the group scale; aw/loc grads are rescaled below
    float go_g[GROUPED ? C::QPW : 1][C::CPL];      // GROUPED: grad_out rows stay in registers across the level tables
    if constexpr (GROUPED) {
#pragma unroll
      for (int pl = 0; pl < C::QPW; ++pl)
        if (pl < npair) Vec16<VT>::load(grad_out + static_cast<int64_t>(pm.pair(p0 + pl)) * D + c * C::CPL, go_g[pl]);
    }
    float g_aw = 0.f, g_x = 0.f, g_y = 0.f;

    const bool g_split = GROUPED && gridDim.y > 1;       // see the forward kernel; grad_loc / grad_aw are then accumulated
    const int g_begin = g_split ? static_cast<int>(blockIdx.y) : 0;
    const int g_end = GROUPED ? (g_split ? g_begin + 1 : G) : 1;
    for (int g = g_begin; g < g_end; ++g) {
      SampleGeom geo;
      int lvl_h = 0, lvl_w = 0;
      Slot sl[4];
#pragma unroll
      for (int cn = 0; cn < 4; ++cn) { sl[cn].off = kInvalidOff; sl[cn].w = 0.f; }
      uint32_t cell_key = kNoMergeKey + 4u * static_cast<uint32_t>(lane & 3);
      if (has_sample) {
        const LevelInfo li = s_lvl[g * L + lvl];
        lvl_h = li.H; lvl_w = li.W;
        geo = make_slot_regs<C::D16>(sl, x, y, a, li, n, m, S, M);
        if (li.H < 32752 && li.W < 32752)
          cell_key = static_cast<uint32_t>(geo.x0 + 8) | (static_cast<uint32_t>(geo.y0 + 8) << 16);
      }
      if (merge == 4) merge_level_slots<4>(sl, cell_key, lane);          // warp-uniform: P == 4 (or 2) and the option is on
      else if (merge == 2) merge_level_slots<2>(sl, cell_key, lane);
      if (has_sample) {
        Slot* dst = my_slots + ps * C::NSLOT + ss;
#pragma unroll
        for (int cn = 0; cn < 4; ++cn) dst[cn * C::LPP] = sl[cn];
      }
      __syncwarp();

#pragma unroll
      for (int pl = 0; pl < C::QPW; ++pl) {
        if (pl < npair) {
          float go[C::CPL];
          if constexpr (GROUPED) {
#pragma unroll
            for (int j = 0; j < C::CPL; ++j) go[j] = go_g[GROUPED ? pl : 0][j];
          } else {
            Vec16<VT>::load(grad_out + static_cast<int64_t>(pm.pair(p0 + pl)) * D + c * C::CPL, go);
          }
          // 2-byte values: a lane owns 8 channels for the dot products, but two 16-byte reductions per lane at a 32-byte lane
          // stride would half-fill every L2 sector.  The scatter only needs grad_out, so for it the lane takes channels
          // [4c, 4c+4) and [4G+4c, 4G+4c+4): each warp-wide RED then covers contiguous 16-byte pieces (bf16 backward 439 -> fp32 speed).
          float go_red[C::CPL == 8 ? 8 : 1];
          if constexpr (C::CPL == 8) {
            const VT* gp = grad_out + static_cast<int64_t>(pm.pair(p0 + pl)) * D;
            const uint2 lo4 = __ldg(reinterpret_cast<const uint2*>(gp + 4 * c));
            const uint2 hi4 = __ldg(reinterpret_cast<const uint2*>(gp + 4 * C::G + 4 * c));
            const uint32_t u4[4] = {lo4.x, lo4.y, hi4.x, hi4.y};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              go_red[2 * i] = __uint_as_float(u4[i] << 16);
              go_red[2 * i + 1] = __uint_as_float(u4[i] & 0xffff0000u);
            }
          }
          const uint4* stream = reinterpret_cast<const uint4*>(my_stream + pl * C::NSLOT);
          // Batches of kBwdBatch corner rows: all gathers of a batch are issued first (predicated, no branches), then the dot
          // products and the reductions (predicated).  The per-lane dot partials of kFold samples are then folded across the
          // G lanes of the corner group by a transposing butterfly (G - 1 shuffles for G samples instead of G log2 G), which
          // leaves the dot of sample i in lane i: one store per lane.  The backward is bound by the SM's load/store + shuffle
          // pipe (gathers, vector reductions, shuffles and shared-memory traffic all queue there), so every shuffle counts.
          constexpr int kFoldT = C::G > kBwdBatch ? C::G : kBwdBatch;
          constexpr bool kTransFold = ((C::G & (C::G - 1)) == 0) && (C::SPG % kFoldT == 0) && (kFoldT % C::G == 0);
          constexpr int kFold = kTransFold ? kFoldT : ((C::SPG % kBwdBatch == 0) ? kBwdBatch : 2);   // samples per fold
          constexpr int kB = (kFold % kBwdBatch == 0) ? kBwdBatch : 2;                               // gathers in flight
          static_assert(C::SPG % kFold == 0 && kFold % kB == 0 && kB % 2 == 0, "batches must tile the slot stream");
#pragma unroll
          for (int f0 = 0; f0 < C::SPG; f0 += kFold) {
            float dot[kFold];
#pragma unroll
            for (int b0 = f0; b0 < f0 + kFold; b0 += kB) {
              uint32_t off[kB];
              float w[kB];
#pragma unroll
              for (int i = 0; i < kB / 2; ++i) {
                const uint4 two = stream[(b0 >> 1) + i];
                off[2 * i] = two.x; w[2 * i] = __uint_as_float(two.y);
                off[2 * i + 1] = two.z; w[2 * i + 1] = __uint_as_float(two.w);
              }
              bool valid[kB];
              float v[kB][C::CPL];
#pragma unroll
              for (int u = 0; u < kB; ++u) {
                valid[u] = active && off[u] != kInvalidOff && !MSDA_DBG_SKIP(8, sgrp * C::SPG + b0 + u);
                if (valid[u]) {
                  Vec16<VT>::load(row_ptr(vlane, off[u]), v[u]);
                } else {
#pragma unroll
                  for (int j = 0; j < C::CPL; ++j) v[u][j] = 0.f;
                }
              }
#pragma unroll
              for (int u = 0; u < kB; ++u) {
                float d = 0.f;
#pragma unroll
                for (int j = 0; j < C::CPL; ++j) d = fmaf(go[j], v[u][j], d);
                dot[b0 - f0 + u] = d;
              }
#pragma unroll
              for (int u = 0; u < kB; ++u) {
                const bool do_red = valid[u] && w[u] != 0.f && !MSDA_DBG_SKIP(4, sgrp * C::SPG + b0 + u);
                if constexpr (C::CPL == 8) {
                  float* gv = grad_value + static_cast<uint64_t>(off[u]) * 8u + 4 * c;      // off counts 16-byte units of 2-byte elements
                  red_add_f32x4_if(do_red, gv, w[u] * go_red[0], w[u] * go_red[1], w[u] * go_red[2], w[u] * go_red[3]);
                  red_add_f32x4_if(do_red, gv + 4 * C::G, w[u] * go_red[4], w[u] * go_red[5], w[u] * go_red[6], w[u] * go_red[7]);
                } else {
                  float* gv = const_cast<float*>(reinterpret_cast<const float*>(
                      reinterpret_cast<const char*>(gvlane) + static_cast<uint64_t>(off[u]) * (16u * sizeof(float) / sizeof(VT))));
#pragma unroll
                  for (int j = 0; j < C::CPL; j += 4)
                    red_add_f32x4_if(do_red, gv + j, w[u] * go[j], w[u] * go[j + 1], w[u] * go[j + 2], w[u] * go[j + 3]);
                }
              }
            }
            if constexpr (kTransFold) {
              // kFold = r * G samples: fold each run of G samples; lane c of the group ends up with sample c of the run
#pragma unroll
              for (int r0 = 0; r0 < kFold; r0 += C::G) {
                float t[C::G];
#pragma unroll
                for (int i = 0; i < C::G; ++i) t[i] = dot[r0 + i];
#pragma unroll
                for (int h = C::G / 2; h >= 1; h >>= 1) {                  // keep the half selected by bit h of the lane
                  const bool up = (c & h) != 0;
#pragma unroll
                  for (int i = 0; i < h; ++i) {
                    const float send = up ? t[i] : t[i + h];
                    const float keep = up ? t[i + h] : t[i];
                    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                  }
                }
                if (active) dot_w[pl * LP + f0 + r0 + c] = t[0];
              }
            } else {
              if constexpr ((C::G & (C::G - 1)) == 0) {
#pragma unroll
                for (int o = C::G / 2; o >= 1; o >>= 1)
#pragma unroll
                  for (int u = 0; u < kFold; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
              } else {                                            // G = 6 or 3: walk down inside the group
#pragma unroll
                for (int u = 0; u < kFold; ++u) {
                  float t = dot[u];
#pragma unroll
                  for (int o = 1; o < C::G; ++o) {
                    const float nb = __shfl_down_sync(0xffffffffu, dot[u], o);
                    if (c + o < C::G) t += nb;
                  }
                  dot[u] = t;
                }
              }
              if (active && c == 0) {
#pragma unroll
                for (int u = 0; u < kFold; ++u) dot_w[pl * LP + f0 + u] = dot[u];
              }
            }
          }
        }
      }
      __syncwarp();

      if (has_sample) {
        float dc[4];                              // per-corner <grad_out, value row>
#pragma unroll
        for (int e = 0; e < 4; ++e) dc[e] = dot_r[e * kDotStride];
        const float hx = 1.f - geo.lx, hy = 1.f - geo.ly;
        g_aw += hy * (hx * dc[0] + geo.lx * dc[1]) + geo.ly * (hx * dc[2] + geo.lx * dc[3]);
        g_x += static_cast<float>(lvl_w) * (hy * (dc[1] - dc[0]) + geo.ly * (dc[3] - dc[2]));
        g_y += static_cast<float>(lvl_h) * (hx * (dc[2] - dc[0]) + geo.lx * (dc[3] - dc[1]));
      }
      __syncwarp();
    }

    if constexpr (kFused) {
      {
        // softmax backward inside the pair's lane segment, chain rule through loc = ref + offsets / scale (+ clamp)
        const float t = a_raw * (scale * g_aw);
        const float tsum = segment_sum<LP>(t);
        if (has_sample) {
          const float inv = 1.f / fz.scale;
          store_pair(grad_loc + 2 * li, a * g_x * inv * mk_x, a * g_y * inv * mk_y);       // d / d raw offsets
          st_from_float(grad_aw + si, t - a_raw * tsum);                                   // d / d logits
        }
        continue;
      }
    }
    if (has_sample) {
      if constexpr (std::is_same<LT, float>::value) {
        if (g_split) {                                       // zero-filled by the host
          atomicAdd(grad_loc + 2 * si, a * g_x);
          atomicAdd(grad_loc + 2 * si + 1, a * g_y);
          atomicAdd(grad_aw + si, scale * g_aw);
        } else {
          store_pair(grad_loc + 2 * si, a * g_x, a * g_y);     // a already carries `scale`
          st_from_float(grad_aw + si, scale * g_aw);
        }
      } else {
        store_pair(grad_loc + 2 * si, a * g_x, a * g_y);
        st_from_float(grad_aw + si, scale * g_aw);
      }
    }
  }
}

}  // namespace msda
