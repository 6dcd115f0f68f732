// Generic multi-scale deformable attention kernels: any head dim D, any L*P, fp32 / bf16 / fp64,
// 64-bit indexing.  They cover what the fast kernels (msda_fast.cuh) do not specialise: the
// reference's gradcheck sweep D in {30,32,64,71,1025,2048,3096} in double (ops/test.py:85-86; the
// reference needs six col2im kernel variants for it, ms_deform_im2col_cuda.cuh:301-920), odd head
// sizes and L*P > 32.  One warp owns one (n, q, m) pair and its lanes stride over the channels, so
// every global access of a warp is one contiguous run of the corner row.
#pragma once

#include "msda_common.cuh"

namespace msda {

template <typename AT>
struct GeomG {
  int x0, y0;
  AT lx, ly;
  bool okx0, okx1, oky0, oky1, any;
};

template <typename AT>
__device__ __forceinline__ GeomG<AT> geom_generic(AT locx, AT locy, int H, int W) {
  GeomG<AT> g;
  const AT x = locx * static_cast<AT>(W) - static_cast<AT>(0.5);
  const AT y = locy * static_cast<AT>(H) - static_cast<AT>(0.5);
  const AT fx = floor(x), fy = floor(y);
  g.lx = x - fx;
  g.ly = y - fy;
  // strict predicate of the reference kernel (ms_deform_im2col_cuda.cuh:288)
  const bool sane = (x > -1) && (x < static_cast<AT>(W)) && (y > -1) && (y < static_cast<AT>(H));
  g.x0 = sane ? static_cast<int>(fx) : -8;
  g.y0 = sane ? static_cast<int>(fy) : -8;
  g.okx0 = g.x0 >= 0 && g.x0 < W;
  g.okx1 = g.x0 + 1 >= 0 && g.x0 + 1 < W;
  g.oky0 = g.y0 >= 0 && g.y0 < H;
  g.oky1 = g.y0 + 1 >= 0 && g.y0 + 1 < H;
  g.any = (g.okx0 || g.okx1) && (g.oky0 || g.oky1);
  return g;
}

template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

template <typename AT>
__device__ __forceinline__ AT warp_sum(AT v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// forward: warp per pair, lanes over channels (loop when D > 32)
template <typename VT, typename LT>
__global__ void __launch_bounds__(kThreads)
msda_fwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                        const LT* __restrict__ aw, VT* __restrict__ out,
                        int S, int M, int D, int L, int Lq, int P, int64_t n_pairs) {
  using AT = typename AccOf<VT>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t row = static_cast<int64_t>(M) * D;
  for (int64_t pair = warp_global; pair < n_pairs; pair += n_warps) {
    const int m = static_cast<int>(pair % M);
    const int64_t n = pair / (static_cast<int64_t>(M) * Lq);
    for (int c0 = 0; c0 < D; c0 += 32) {
      const int c = c0 + lane;
      const bool cok = c < D;
      AT acc = 0;
      for (int l = 0; l < L; ++l) {
        const bool fits = level_fits(shapes[2 * l], shapes[2 * l + 1], level_start[l], S);      // see msda_common.cuh
        const int H = fits ? static_cast<int>(shapes[2 * l]) : 0, W = fits ? static_cast<int>(shapes[2 * l + 1]) : 0;
        const VT* vl = value + (n * S + (fits ? level_start[l] : 0)) * row + static_cast<int64_t>(m) * D;
        for (int p = 0; p < P; ++p) {
          const int64_t si = (pair * L + l) * P + p;
          const AT a = static_cast<AT>(ld_as_float(aw + si));
          const GeomG<AT> g = geom_generic<AT>(static_cast<AT>(ld_as_float(loc + 2 * si)),
                                               static_cast<AT>(ld_as_float(loc + 2 * si + 1)), H, W);
          if (!g.any || !cok) continue;
          const VT* p00 = vl + (static_cast<int64_t>(g.y0) * W + g.x0) * row + c;
          const AT hx = 1 - g.lx, hy = 1 - g.ly;
          AT v = 0;
          if (g.oky0 && g.okx0) v += hx * hy * static_cast<AT>(ld_as_float(p00));
          if (g.oky0 && g.okx1) v += g.lx * hy * static_cast<AT>(ld_as_float(p00 + row));
          if (g.oky1 && g.okx0) v += hx * g.ly * static_cast<AT>(ld_as_float(p00 + static_cast<int64_t>(W) * row));
          if (g.oky1 && g.okx1) v += g.lx * g.ly * static_cast<AT>(ld_as_float(p00 + static_cast<int64_t>(W) * row + row));
          acc += a * v;
        }
      }
      if (cok) st_from_float(out + pair * D + c, acc);
    }
  }
}

// backward: warp per pair; grad_value scattered with scalar atomics into an accumulation image of
// type GT (float for fp32/bf16 tensors, double for fp64), the three per-sample sums folded with
// warp shuffles.
template <typename VT, typename LT, typename GT>
__global__ void __launch_bounds__(kThreads)
msda_bwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ level_start, const LT* __restrict__ loc,
                        const LT* __restrict__ aw, const VT* __restrict__ grad_out,
                        GT* __restrict__ grad_value, LT* __restrict__ grad_loc, LT* __restrict__ grad_aw,
                        int S, int M, int D, int L, int Lq, int P, int64_t n_pairs) {
  using AT = typename AccOf<VT>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t row = static_cast<int64_t>(M) * D;
  for (int64_t pair = warp_global; pair < n_pairs; pair += n_warps) {
    const int m = static_cast<int>(pair % M);
    const int64_t n = pair / (static_cast<int64_t>(M) * Lq);
    const VT* go = grad_out + pair * D;
    for (int l = 0; l < L; ++l) {
      const bool fits = level_fits(shapes[2 * l], shapes[2 * l + 1], level_start[l], S);
      const int H = fits ? static_cast<int>(shapes[2 * l]) : 0, W = fits ? static_cast<int>(shapes[2 * l + 1]) : 0;
      const int64_t lbase = (n * S + (fits ? level_start[l] : 0)) * row + static_cast<int64_t>(m) * D;
      for (int p = 0; p < P; ++p) {
        const int64_t si = (pair * L + l) * P + p;
        const AT a = static_cast<AT>(ld_as_float(aw + si));
        const GeomG<AT> g = geom_generic<AT>(static_cast<AT>(ld_as_float(loc + 2 * si)),
                                             static_cast<AT>(ld_as_float(loc + 2 * si + 1)), H, W);
        AT s_aw = 0, s_x = 0, s_y = 0;
        if (g.any) {
          const AT hx = 1 - g.lx, hy = 1 - g.ly;
          const int64_t o00 = lbase + (static_cast<int64_t>(g.y0) * W + g.x0) * row;
          const int64_t o01 = o00 + row, o10 = o00 + static_cast<int64_t>(W) * row, o11 = o10 + row;
          for (int c = lane; c < D; c += 32) {
            const AT gch = static_cast<AT>(ld_as_float(go + c));
            const AT ga = gch * a;
            AT v00 = 0, v01 = 0, v10 = 0, v11 = 0;
            if (g.oky0 && g.okx0) { v00 = static_cast<AT>(ld_as_float(value + o00 + c)); atomicAdd(grad_value + o00 + c, static_cast<GT>(hx * hy * ga)); }
            if (g.oky0 && g.okx1) { v01 = static_cast<AT>(ld_as_float(value + o01 + c)); atomicAdd(grad_value + o01 + c, static_cast<GT>(g.lx * hy * ga)); }
            if (g.oky1 && g.okx0) { v10 = static_cast<AT>(ld_as_float(value + o10 + c)); atomicAdd(grad_value + o10 + c, static_cast<GT>(hx * g.ly * ga)); }
            if (g.oky1 && g.okx1) { v11 = static_cast<AT>(ld_as_float(value + o11 + c)); atomicAdd(grad_value + o11 + c, static_cast<GT>(g.lx * g.ly * ga)); }
            s_aw += gch * (hy * (hx * v00 + g.lx * v01) + g.ly * (hx * v10 + g.lx * v11));
            s_x += gch * (hy * (v01 - v00) + g.ly * (v11 - v10));
            s_y += gch * (hx * (v10 - v00) + g.lx * (v11 - v01));
          }
        }
        s_aw = warp_sum(s_aw);
        s_x = warp_sum(s_x);
        s_y = warp_sum(s_y);
        if (lane == 0) {
          st_from_float(grad_aw + si, s_aw);
          st_from_float(grad_loc + 2 * si, a * static_cast<AT>(W) * s_x);
          st_from_float(grad_loc + 2 * si + 1, a * static_cast<AT>(H) * s_y);
        }
      }
    }
  }
}

}  // namespace msda
