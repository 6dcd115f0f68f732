// EXPERIMENT (selected only with option bwd_variant = 5): backward kernel with on-chip pre-reduction of grad_value for
// the coarse pyramid levels.  Parity-green (the encoder parity tests pass on it) but MEASURED SLOWER than
// msda_bwd_fast2_kernel on B200 (encoder R50_360, tools/kernel_bench.py, profiles/r01o_kernel_bench_quick.json):
//   fast2 244.7 us | tile: level 3 on chip 242.8 us | levels 2+3 on chip, 64 queries/CTA 275.5 us, 32: 269 us, 128: 310 us
// i.e. 4 native ATOMS.ADD per lane and corner cost more than the one 16-byte L2 reduction they replace, and the lower
// occupancy (64 registers, 40 KB tile) hurts the gathers.  Kept as the documented negative result for this design point.
//
// Why: msda_bwd_fast2_kernel is bound by L2 (profiles/r01l_ncu_full.md: 37.7 M reduction sectors per encoder call,
// half of them crossing the die-to-die fabric).  Half of all samples land on the two coarsest levels, whose value rows
// are few (R50_ovis_360: 240 + 60 positions per head and frame) but are hit ~340 and ~1360 times each.  Floating-point
// shared-memory atomics are CAS loops on sm_100a (ATOMS.CAST.SPIN), but 32-bit INTEGER shared atomics are native
// (ATOMS.ADD), so the coarse levels are accumulated in shared memory in per-CTA fixed point:
//
//   * a CTA owns one (frame n, head m) and a run of `q_per_cta` queries; its tile holds every row of the selected
//     levels for that head: int32 [rows][33] (pitch 33 words: the four corner rows of a sample fall in different banks);
//   * scale = 2.147e9 / (max|grad_out| * sum|attention weight|) over exactly the CTA's own pairs and tile-level samples.
//     Every contribution is w * go with |w| <= |aw| (bilinear weights are <= 1), so |sum| <= max|go| * sum|aw|: the
//     accumulators cannot overflow, and the resolution adapts to the data (2^-31 of that bound per contribution,
//     i.e. ~1e-8 relative to max|go| for 64 queries -- below fp32 rounding of the atomics it replaces);
//   * at the end the tile is converted back and leaves with ONE vector reduction per row and 4 channels -- per CTA
//     `rows` instead of 32 * q_per_cta reductions for those levels.
// Finer levels keep the direct red.global.add.v4.f32 path.  Everything else (slot records, corner-group gathers,
// shuffle-folded dots, phase 3) is msda_bwd_fast2_kernel's.  fp32, L = P = 4 (MDQE's spatial attention), D in {32, 24}.
#pragma once

#include "msda_fast2.cuh"

namespace msda {

constexpr int kTilePitch = 33;

template <int D>
__global__ void __launch_bounds__(kThreads, 4)
msda_bwd_tile_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ level_start, const float* __restrict__ loc,
                     const float* __restrict__ aw, const float* __restrict__ grad_out,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_aw,
                     int S, int M, int Lq, int q_per_cta, int tile_rows_max) {
  constexpr int L = 4, P = 4, LP = 16;
  using C = Cfg2<float, D, LP>;
  static_assert(C::NSG == 1 && C::QPW == 2, "fp32 configuration expected");
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  int* tile = reinterpret_cast<int*>(dyn_smem);                       // [tile_rows][kTilePitch]
  __shared__ LevelInfo s_lvl[L];
  __shared__ int s_tile_base[L + 1];
  __shared__ __align__(16) Slot s_slot[kWarpsPerCta][C::QPW * C::NSLOT];
  __shared__ float s_dot[kWarpsPerCta][4 * 33];
  __shared__ float s_red[2][kWarpsPerCta];
  __shared__ float s_scale[2];
  __shared__ uint32_t s_mask;

  const int n = blockIdx.z, m = blockIdx.y;
  const int q_begin = blockIdx.x * q_per_cta;
  const int q_end = min(Lq, q_begin + q_per_cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  stage_levels(s_lvl, shapes, level_start, L);
  __syncthreads();
  if (threadIdx.x == 0) {
    // levels kept on chip: greedily the smallest ones while their rows fit the shared-memory budget (the shapes live
    // on the device, as in the reference API, so the choice is made here rather than on the host)
    uint32_t mask = 0;
    int used = 0;
    for (int round = 0; round < L; ++round) {
      int best = -1, best_rows = 0x7fffffff;
      for (int l = 0; l < L; ++l) {
        const int rows = s_lvl[l].H * s_lvl[l].W;
        if (!((mask >> l) & 1u) && rows < best_rows) { best = l; best_rows = rows; }
      }
      if (best < 0 || used + best_rows > tile_rows_max) break;
      mask |= 1u << best;
      used += best_rows;
    }
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      s_tile_base[l] = acc;
      if ((mask >> l) & 1u) acc += s_lvl[l].H * s_lvl[l].W;
    }
    s_tile_base[L] = acc;
    s_mask = mask;
  }
  __syncthreads();
  const uint32_t tile_mask = s_mask;
  const int tile_rows = s_tile_base[L];
  for (int i = threadIdx.x; i < tile_rows * kTilePitch; i += kThreads) tile[i] = 0;

  // ---- per-CTA fixed-point scale: max |grad_out| over the CTA's pairs, sum |aw| over its tile-level samples
  {
    float gmax = 0.f, asum = 0.f;
    const int nq = q_end - q_begin;
    for (int i = threadIdx.x; i < nq * (D / 4); i += kThreads) {
      const int qi = i / (D / 4), c4 = i - qi * (D / 4);
      const float4 g = __ldg(reinterpret_cast<const float4*>(grad_out + ((static_cast<int64_t>(n) * Lq + q_begin + qi) * M + m) * D) + c4);
      gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
    }
    for (int i = threadIdx.x; i < nq * LP; i += kThreads) {
      const int qi = i / LP, s = i - qi * LP;
      if ((tile_mask >> (s / P)) & 1u) asum += fabsf(__ldg(aw + ((static_cast<int64_t>(n) * Lq + q_begin + qi) * M + m) * LP + s));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
      asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    if (lane == 0) { s_red[0][warp] = gmax; s_red[1][warp] = asum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float g = 0.f, a = 0.f;
      for (int w = 0; w < kWarpsPerCta; ++w) { g = fmaxf(g, s_red[0][w]); a += s_red[1][w]; }
      const float bound = g * a;
      // non-finite gradients: fall back to scale 0 (contributions vanish) rather than poisoning the integer tile
      const bool ok = bound > 0.f && bound < 3.0e38f;
      s_scale[0] = ok ? 2.147e9f / bound : 0.f;
      s_scale[1] = ok ? bound / 2.147e9f : 0.f;
    }
    __syncthreads();
  }
  const float fx_scale = s_scale[0], fx_inv = s_scale[1];

  const int grp = lane / C::G, c = lane - grp * C::G;
  const bool active = C::kAllLanes || grp < C::NG;
  const int corner = grp & 3;
  Slot* my_slots = s_slot[warp];
  const Slot* my_stream = my_slots + corner * C::LPP;
  const float* vlane = value + c * C::CPL;
  float* gvlane = grad_value + c * C::CPL;
  float* dot_w = s_dot[warp] + corner * 33;
  const float* dot_r = s_dot[warp] + lane;
  const int ps = lane >> 4, ss = lane & 15;           // phase-1 role: which of the two pairs, which sample
  const int lvl = ss >> 2;
  const uint32_t row16 = static_cast<uint32_t>(M) * C::D16;
  // 16-byte-unit offset of (n, level l, position 0, head m) and tile row base, per level
  uint32_t lvl_base16[L];
  int tbase[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    lvl_base16[l] = (static_cast<uint32_t>(n) * S + s_lvl[l].start) * row16 + static_cast<uint32_t>(m) * C::D16;
    tbase[l] = s_tile_base[l];
  }

  for (int q0 = q_begin + warp * 2; q0 < q_end; q0 += kWarpsPerCta * 2) {
    const int npair = min(2, q_end - q0);
    const bool has_sample = lane < npair * LP;
    float x = 0.f, y = 0.f, a = 0.f;
    SampleGeom geo;
    int lvl_h = 0, lvl_w = 0;
    int64_t si = 0;
    if (has_sample) {
      const int64_t pair = (static_cast<int64_t>(n) * Lq + q0 + ps) * M + m;
      si = pair * LP + ss;
      load_loc_aw<float>(loc, aw, si, x, y, a);
      const LevelInfo li = s_lvl[lvl];
      lvl_h = li.H; lvl_w = li.W;
      // slot records hold the position index inside the level (the row offsets are rebuilt per level in phase 2)
      geo = sample_geom(x, y, li.H, li.W);
      const int cell00 = geo.y0 * li.W + geo.x0;
      const float hx = 1.f - geo.lx, hy = 1.f - geo.ly;
      const bool v00 = geo.oky0 && geo.okx0, v01 = geo.oky0 && geo.okx1, v10 = geo.oky1 && geo.okx0, v11 = geo.oky1 && geo.okx1;
      Slot* dst = my_slots + ps * C::NSLOT;
      Slot e;
      e.off = v00 ? static_cast<uint32_t>(cell00) : kInvalidOff;             e.w = v00 ? hx * hy * a : 0.f;          dst[0 * C::LPP + ss] = e;
      e.off = v01 ? static_cast<uint32_t>(cell00 + 1) : kInvalidOff;         e.w = v01 ? geo.lx * hy * a : 0.f;      dst[1 * C::LPP + ss] = e;
      e.off = v10 ? static_cast<uint32_t>(cell00 + li.W) : kInvalidOff;      e.w = v10 ? hx * geo.ly * a : 0.f;      dst[2 * C::LPP + ss] = e;
      e.off = v11 ? static_cast<uint32_t>(cell00 + li.W + 1) : kInvalidOff;  e.w = v11 ? geo.lx * geo.ly * a : 0.f;  dst[3 * C::LPP + ss] = e;
    }
    __syncwarp();

#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      if (pl < npair) {
        float go[C::CPL], gs[C::CPL];
        Vec16<float>::load(grad_out + ((static_cast<int64_t>(n) * Lq + q0 + pl) * M + m) * D + c * C::CPL, go);
#pragma unroll
        for (int j = 0; j < C::CPL; ++j) gs[j] = go[j] * fx_scale;
        const uint4* stream = reinterpret_cast<const uint4*>(my_stream + pl * C::NSLOT);
#pragma unroll
        for (int it = 0; it < LP / 2; ++it) {
          const uint4 two = stream[it];
          const uint32_t cell[2] = {two.x, two.z};
          const float w[2] = {__uint_as_float(two.y), __uint_as_float(two.w)};
          const int l = (2 * it) / P;                              // both samples of the pair share the level (P = 4)
          const bool in_tile = (tile_mask >> l) & 1u;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float dot = 0.f;
            if (active && cell[u] != kInvalidOff) {
              const uint32_t off16 = lvl_base16[l] + cell[u] * row16;
              float v[C::CPL];
              Vec16<float>::load(row_ptr(vlane, off16), v);
#pragma unroll
              for (int j = 0; j < C::CPL; ++j) dot = fmaf(go[j], v[j], dot);
              if (in_tile) {
                int* t = tile + (tbase[l] + static_cast<int>(cell[u])) * kTilePitch + c * C::CPL;
#pragma unroll
                for (int j = 0; j < C::CPL; ++j) atomicAdd(t + j, __float2int_rn(w[u] * gs[j]));
              } else {
                float* gv = const_cast<float*>(reinterpret_cast<const float*>(reinterpret_cast<const char*>(gvlane) + static_cast<uint64_t>(off16) * 16u));
                red_add_f32x4(gv, w[u] * go[0], w[u] * go[1], w[u] * go[2], w[u] * go[3]);
              }
            }
            if constexpr ((C::G & (C::G - 1)) == 0) {
#pragma unroll
              for (int o = C::G / 2; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            } else {
              float t2 = dot;
#pragma unroll
              for (int o = 1; o < C::G; ++o) {
                const float nb = __shfl_down_sync(0xffffffffu, dot, o);
                if (c + o < C::G) t2 += nb;
              }
              dot = t2;
            }
            if (active && c == 0) dot_w[pl * LP + 2 * it + u] = dot;
          }
        }
      }
    }
    __syncwarp();

    if (has_sample) {
      float dc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) dc[e] = dot_r[e * 33];
      const float hx = 1.f - geo.lx, hy = 1.f - geo.ly;
      const float g_aw = hy * (hx * dc[0] + geo.lx * dc[1]) + geo.ly * (hx * dc[2] + geo.lx * dc[3]);
      const float g_x = a * static_cast<float>(lvl_w) * (hy * (dc[1] - dc[0]) + geo.ly * (dc[3] - dc[2]));
      const float g_y = a * static_cast<float>(lvl_h) * (hx * (dc[2] - dc[0]) + geo.lx * (dc[3] - dc[1]));
      store_pair(grad_loc + 2 * si, g_x, g_y);
      grad_aw[si] = g_aw;
    }
    __syncwarp();
  }

  // ---- flush the tile: fixed point -> fp32, one vector reduction per row and 4 channels
  __syncthreads();
  for (int i = threadIdx.x; i < tile_rows * (D / 4); i += kThreads) {
    const int r = i / (D / 4), c4 = i - r * (D / 4);
    const int* t = tile + r * kTilePitch + c4 * 4;
    const int i0 = t[0], i1 = t[1], i2 = t[2], i3 = t[3];
    if ((i0 | i1 | i2 | i3) == 0) continue;
    int l = 0;                                   // the selected level whose row range contains r
#pragma unroll
    for (int k = 0; k < L; ++k)
      if (((tile_mask >> k) & 1u) && r >= s_tile_base[k]) l = k;
    const int cell = r - s_tile_base[l];
    float* dst = grad_value + ((static_cast<int64_t>(n) * S + s_lvl[l].start + cell) * M + m) * D + c4 * 4;
    red_add_f32x4(dst, static_cast<float>(i0) * fx_inv, static_cast<float>(i1) * fx_inv, static_cast<float>(i2) * fx_inv,
                  static_cast<float>(i3) * fx_inv);
  }
}

}  // namespace msda
