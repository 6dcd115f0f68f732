// Fast multi-scale deformable attention kernels for the MDQE head sizes (D = 32: R50, D = 24: Swin-L).
//
// Replaces ms_deformable_im2col_gpu_kernel (/root/reference/mdqe/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:237-299) and the D<=32 shared-memory col2im kernels (:301-403, :513-615).
//
// Work decomposition ("pair" = one (batch n, query q, head m) triple, `LP = L*P` samples each):
//   * a CTA owns `chunk_pairs` consecutive pairs (consecutive queries x all heads), its 8 warps
//     walk them with a fixed warp->head assignment, so one warp streams over neighbouring queries of
//     the same head and the bilinear corner rows it needs are the ones it just pulled into L1;
//   * phase 1 (one lane per sample): coalesced load of loc/aw for 32/LP pairs at once, bilinear
//     geometry, and a 4-entry "slot" record per sample {row offset, corner weight * attention} in
//     shared memory -- loc/aw/shape words are read once per sample instead of once per channel;
//   * phase 2 (one lane group per bilinear corner): a group of G = D*sizeof(VT)/16 lanes fetches one
//     corner's channel row with a single 16-byte load per lane, so each warp instruction moves 4 (fp32)
//     or 8 (bf16) complete corner rows; partial sums stay in registers and are folded across the
//     groups with shuffles once per pair;
//   * backward phase 2 additionally scatters grad_value with 16-byte vector reductions
//     (red.global.add.v4.f32 -> one L2 atomic per 4 channels instead of 4) and leaves per-lane
//     <grad_out, corner> dot products in shared memory; phase 3 (one lane per sample again) folds
//     them into grad_attn_weight and grad_sampling_loc and writes both coalesced.
#pragma once

#include "msda_common.cuh"

namespace msda {

template <typename VT, int D>
struct FastCfg {
  static constexpr int CPL = 16 / static_cast<int>(sizeof(VT));   // channels per lane
  static constexpr int G = D / CPL;                                // lanes per corner row
  static constexpr int NG = (32 / G >= 8) ? 8 : 4;                 // corner groups per warp
  static constexpr int ROW = 4 * G;                                // dot partials per sample
  static_assert(D % CPL == 0, "head dim must be a multiple of the 16-byte vector");
  static_assert(G * NG <= 32, "groups must fit a warp");
};

template <typename LT>
__device__ __forceinline__ void load_loc_aw(const LT* __restrict__ loc, const LT* __restrict__ aw, int64_t si,
                                            float& x, float& y, float& a);
template <>
__device__ __forceinline__ void load_loc_aw<float>(const float* __restrict__ loc, const float* __restrict__ aw,
                                                   int64_t si, float& x, float& y, float& a) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(loc) + si);
  x = t.x; y = t.y; a = __ldg(aw + si);
}
template <>
__device__ __forceinline__ void load_loc_aw<__nv_bfloat16>(const __nv_bfloat16* __restrict__ loc,
                                                           const __nv_bfloat16* __restrict__ aw, int64_t si,
                                                           float& x, float& y, float& a) {
  const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(loc) + si);
  x = __uint_as_float(t << 16); y = __uint_as_float(t & 0xffff0000u);
  a = __bfloat162float(aw[si]);
}

// two-index form: loc is addressed in (x, y) pairs by `li`, aw by `ai`.  They differ only for the joint query projection of the
// fused prologue (fp32; FusedArgs::row_stride), where offsets and logits are column ranges of one [N*Lq, row_stride] matrix.
template <typename LT>
__device__ __forceinline__ void load_loc_aw2(const LT* __restrict__ loc, const LT* __restrict__ aw, int64_t li, int64_t ai,
                                             float& x, float& y, float& a) {
  load_loc_aw<LT>(loc, aw, li, x, y, a);
}
template <>
__device__ __forceinline__ void load_loc_aw2<float>(const float* __restrict__ loc, const float* __restrict__ aw, int64_t li,
                                                    int64_t ai, float& x, float& y, float& a) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(loc) + li);
  x = t.x; y = t.y; a = __ldg(aw + ai);
}

__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

constexpr uint32_t kInvalidOff = 0xffffffffu;

// Phase 1 for one sample: geometry + the four slot records (corner order: (y0,x0) (y0,x1) (y1,x0)
// (y1,x1)).  Returns the geometry (the backward needs it again in phase 3).  Out-of-range corners
// are marked with kInvalidOff; the forward additionally skips corners whose weight is exactly 0,
// the backward must not (grad_attn_weight does not depend on the attention weight itself).
template <int D>
__device__ __forceinline__ SampleGeom make_slots(Slot* dst, float locx, float locy, float a, const LevelInfo li,
                                                 int n, int m, int S, int M) {
  const SampleGeom g = sample_geom(locx, locy, li.H, li.W);
  const uint32_t row = static_cast<uint32_t>(M) * D;
  const uint32_t base = (static_cast<uint32_t>(n) * S + li.start) * row + static_cast<uint32_t>(m) * D;
  const uint32_t o00 = base + static_cast<uint32_t>(g.y0 * li.W + g.x0) * row;   // wraps harmlessly when invalid
  const float hx = 1.f - g.lx, hy = 1.f - g.ly;
  uint4 lo, hi;
  lo.x = (g.oky0 && g.okx0) ? o00 : kInvalidOff;                      lo.y = __float_as_uint(hx * hy * a);
  lo.z = (g.oky0 && g.okx1) ? o00 + row : kInvalidOff;                lo.w = __float_as_uint(g.lx * hy * a);
  hi.x = (g.oky1 && g.okx0) ? o00 + li.W * row : kInvalidOff;         hi.y = __float_as_uint(hx * g.ly * a);
  hi.z = (g.oky1 && g.okx1) ? o00 + li.W * row + row : kInvalidOff;   hi.w = __float_as_uint(g.lx * g.ly * a);
  reinterpret_cast<uint4*>(dst)[0] = lo;
  reinterpret_cast<uint4*>(dst)[1] = hi;
  return g;
}

// ------------------------------------------------------------------------------------------ forward
template <typename VT, typename LT, int D>
__global__ void __launch_bounds__(kThreads)
msda_fwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                     const LT* __restrict__ aw, VT* __restrict__ out,
                     int S, int M, int L, int Lq, int P, int64_t n_pairs, int chunk_pairs) {
  using C = FastCfg<VT, D>;
  __shared__ LevelInfo s_lvl[kMaxLevels];
  __shared__ __align__(16) Slot s_slot[kWarpsPerCta][128];

  stage_levels(s_lvl, shapes, level_start, L, S);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int LP = L * P;                       // <= 32 (host-checked)
  const int qpw = 32 / LP;                    // pairs a warp prepares per round
  const int grp = lane / C::G, c = lane - grp * C::G;
  const bool active = grp < C::NG;
  const int nslot = LP * 4;
  const int iters = (nslot + C::NG - 1) / C::NG;
  Slot* my_slots = s_slot[warp];

  const int64_t chunk_begin = static_cast<int64_t>(blockIdx.x) * chunk_pairs;
  const int64_t chunk_end = min(n_pairs, chunk_begin + chunk_pairs);

  for (int64_t p0 = chunk_begin + warp * qpw; p0 < chunk_end; p0 += kWarpsPerCta * qpw) {
    const int npair = static_cast<int>(min(static_cast<int64_t>(qpw), chunk_end - p0));
    if (lane < npair * LP) {
      const int pl = lane / LP, s = lane - pl * LP;
      const int64_t pair = p0 + pl;
      const int m = static_cast<int>(pair % M);
      const int n = static_cast<int>(pair / (static_cast<int64_t>(M) * Lq));
      float x, y, a;
      load_loc_aw<LT>(loc, aw, p0 * LP + lane, x, y, a);
      make_slots<D>(my_slots + lane * 4, x, y, a, s_lvl[s / P], n, m, S, M);
    }
    __syncwarp();

    for (int pl = 0; pl < npair; ++pl) {
      float acc[C::CPL];
#pragma unroll
      for (int j = 0; j < C::CPL; ++j) acc[j] = 0.f;
      const Slot* ps = my_slots + pl * nslot;
#pragma unroll 8
      for (int it = 0; it < iters; ++it) {
        const int sl = it * C::NG + grp;
        Slot e;
        e.off = kInvalidOff; e.w = 0.f;
        if (active && sl < nslot) e = ps[sl];
        if (e.off != kInvalidOff && e.w != 0.f) {
          float v[C::CPL];
          Vec16<VT>::load(value + e.off + c * C::CPL, v);
#pragma unroll
          for (int j = 0; j < C::CPL; ++j) acc[j] = fmaf(e.w, v[j], acc[j]);
        }
      }
#pragma unroll
      for (int k = C::NG / 2; k >= 1; k >>= 1) {
#pragma unroll
        for (int j = 0; j < C::CPL; ++j) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], k * C::G);
      }
      if (lane < C::G) Vec16<VT>::store(out + (p0 + pl) * D + lane * C::CPL, acc);
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------- backward
// GVT is the type of the grad_value accumulation image: always float here (for bf16 tensors the
// host passes the fp32 workspace and converts afterwards).
template <typename VT, typename LT, int D>
__global__ void __launch_bounds__(kThreads)
msda_bwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                     const LT* __restrict__ aw, const VT* __restrict__ grad_out,
                     float* __restrict__ grad_value, LT* __restrict__ grad_loc, LT* __restrict__ grad_aw,
                     int S, int M, int L, int Lq, int P, int64_t n_pairs, int chunk_pairs) {
  using C = FastCfg<VT, D>;
  __shared__ LevelInfo s_lvl[kMaxLevels];
  __shared__ __align__(16) Slot s_slot[kWarpsPerCta][128];
  __shared__ __align__(16) float s_dot[kWarpsPerCta][32 * C::ROW];

  stage_levels(s_lvl, shapes, level_start, L, S);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int LP = L * P;
  const int qpw = 32 / LP;
  const int grp = lane / C::G, c = lane - grp * C::G;
  const bool active = grp < C::NG;
  const int nslot = LP * 4;
  const int iters = (nslot + C::NG - 1) / C::NG;
  Slot* my_slots = s_slot[warp];
  float* my_dot = s_dot[warp];

  const int64_t chunk_begin = static_cast<int64_t>(blockIdx.x) * chunk_pairs;
  const int64_t chunk_end = min(n_pairs, chunk_begin + chunk_pairs);

  for (int64_t p0 = chunk_begin + warp * qpw; p0 < chunk_end; p0 += kWarpsPerCta * qpw) {
    const int npair = static_cast<int>(min(static_cast<int64_t>(qpw), chunk_end - p0));
    const bool has_sample = lane < npair * LP;
    SampleGeom geo;
    float a = 0.f;
    int lvl_h = 0, lvl_w = 0;
    if (has_sample) {
      const int pl = lane / LP, s = lane - pl * LP;
      const int64_t pair = p0 + pl;
      const int m = static_cast<int>(pair % M);
      const int n = static_cast<int>(pair / (static_cast<int64_t>(M) * Lq));
      float x, y;
      load_loc_aw<LT>(loc, aw, p0 * LP + lane, x, y, a);
      const LevelInfo li = s_lvl[s / P];
      lvl_h = li.H; lvl_w = li.W;
      geo = make_slots<D>(my_slots + lane * 4, x, y, a, li, n, m, S, M);
    }
    __syncwarp();

    for (int pl = 0; pl < npair; ++pl) {
      float go[C::CPL];
      Vec16<VT>::load(grad_out + (p0 + pl) * D + c * C::CPL, go);   // idle lanes read a valid address too
      const Slot* ps = my_slots + pl * nslot;
#pragma unroll 4
      for (int it = 0; it < iters; ++it) {
        const int sl = it * C::NG + grp;
        if (active && sl < nslot) {
          const Slot e = ps[sl];
          float dot = 0.f;
          if (e.off != kInvalidOff) {
            float v[C::CPL];
            Vec16<VT>::load(value + e.off + c * C::CPL, v);
#pragma unroll
            for (int j = 0; j < C::CPL; ++j) dot = fmaf(go[j], v[j], dot);
            float* gv = grad_value + e.off + c * C::CPL;
#pragma unroll
            for (int j = 0; j < C::CPL; j += 4)
              red_add_f32x4(gv + j, e.w * go[j], e.w * go[j + 1], e.w * go[j + 2], e.w * go[j + 3]);
          }
          // dot partial of (sample, corner, lane-in-group); 16-byte chunks rotated by the sample
          // index so that phase 3 (lane = sample) reads them without bank conflicts.
          const int smp = pl * LP + (sl >> 2);
          const int e_idx = (sl & 3) * C::G + c;
          const int chunk = ((e_idx >> 2) + smp) % C::G;
          my_dot[smp * C::ROW + chunk * 4 + (e_idx & 3)] = dot;
        }
      }
    }
    __syncwarp();

    if (has_sample) {
      float dc[4] = {0.f, 0.f, 0.f, 0.f};     // per-corner <grad_out, value row>
      const float4* row = reinterpret_cast<const float4*>(my_dot + lane * C::ROW);
#pragma unroll
      for (int k = 0; k < C::G; ++k) {
        const float4 t = row[(k + lane) % C::G];
        dc[(4 * k + 0) / C::G] += t.x;
        dc[(4 * k + 1) / C::G] += t.y;
        dc[(4 * k + 2) / C::G] += t.z;
        dc[(4 * k + 3) / C::G] += t.w;
      }
      const float hx = 1.f - geo.lx, hy = 1.f - geo.ly;
      const float g_aw = hy * (hx * dc[0] + geo.lx * dc[1]) + geo.ly * (hx * dc[2] + geo.lx * dc[3]);
      const float g_x = a * static_cast<float>(lvl_w) * (hy * (dc[1] - dc[0]) + geo.ly * (dc[3] - dc[2]));
      const float g_y = a * static_cast<float>(lvl_h) * (hx * (dc[2] - dc[0]) + geo.lx * (dc[3] - dc[1]));
      const int64_t si = p0 * LP + lane;
      store_pair(grad_loc + 2 * si, g_x, g_y);
      st_from_float(grad_aw + si, g_aw);
    }
    __syncwarp();
  }
}

// fp32 accumulation image -> bf16 grad_value
static __global__ void __launch_bounds__(256)
cvt_f32_to_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = src[i];
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

}  // namespace msda
