// Callers either side of the hot path, sm_100a SIMT kernels behind the C ABI of include/msda_b200.h.
// Paths relative to /root/reference:
//   mask_match_cost          mdqe/models/matcher.py:182-200 with batch_sigmoid_ce_loss (:36-61) and batch_dice_loss (:11-28):
//                            the Hungarian matcher's mask costs, fused with the mask contraction -- out_masks [Q, T*H*W]
//                            (48 MB per clip at R50_ovis_360) is never written, proto and the targets are read once;
//   mask_nms_siou            mdqe/mdqe.py:386-399: soft-IoU matrix of inference_clip (frame stride 2 for clips of 5+ frames,
//                            nearest 0.5x downsampling, sigmoid, threshold, Q x Q product) in one pass over mask_pred;
//   aligned_bilinear_sigmoid mdqe/util/misc.py:485-507 + mdqe/mdqe.py:357: the 4x mask upsampling of the inference output;
//   query_init_sample_*      mdqe/models/transformer_dec.py:170-179: per-level F.grid_sample (bilinear, border padding,
//                            align_corners=False) of the encoder memory at the selected query points, mean over levels.
// All of these are memory- or latency-bound at MDQE's sizes; none is a contraction large enough for the tensor cores
// (the second product of the matcher cost has 5-20 target columns).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "msda_internal.h"

namespace msda {
namespace {

constexpr int kTC = 32;        // plane columns per iteration
constexpr int kPad = 36;       // row stride (floats) of the shared tiles: keeps float4 alignment, rows 4 banks apart
constexpr int kThreads = 256;

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------ matcher mask costs
// One launch handles up to 207 queries and 15 targets (68 KB of shared memory: three CTAs per SM).  Per 32 plane columns:
//   stage 1  x = coeff . proto for 4 x 4 (query, column) tiles, sigmoid(x) to shared memory, softplus(x) summed per query;
//   stage 2  sigmoid (rows) x targets (columns), 4 x 4 tile per thread.  Row Q of the sigmoid tile is all ones (gives
//            sum_c tgt[g,c]) and target G is all ones (gives sum_c sigmoid[q,c]): both row sums fall out of the same product;
//   stage 3  PT[k,g] += sum_c proto[k,c] tgt[g,c].  The BCE term needs sum_c x[q,c] tgt[g,c] = sum_k coeff[q,k] PT[k,g]: by
//            associativity the big product over the plane is 32 x G instead of Q x G, and x never goes to shared memory.
constexpr int kMcMaxQ = 208, kMcMaxG = 16, kMcK = 32;
constexpr int kMcWsFloats = kMcMaxQ * kMcMaxG + kMcK * kMcMaxG + kMcMaxQ;      // sigmoid x targets, PT, softplus row sums
constexpr size_t kMcSmem = static_cast<size_t>(kMcMaxQ + kMcK + kMcMaxG + kMcMaxQ) * kPad * sizeof(float);

__global__ void __launch_bounds__(kThreads, 3)
match_cost_kernel(const float* __restrict__ coeff, const float* __restrict__ proto, const float* __restrict__ tgt, int Q, int K, int G,
                  int64_t N, int64_t cols_per_cta, float* __restrict__ ws) {
  extern __shared__ float4 smem4[];
  float* s_coeff = reinterpret_cast<float*>(smem4);          // [kMcMaxQ][kPad]  coeff[q][k], zero padded
  float* s_proto = s_coeff + kMcMaxQ * kPad;                 // [32][kPad]       proto[k][c]
  float* s_tgt = s_proto + kMcK * kPad;                      // [kMcMaxG][kPad]  tgt[g][c], row G = 1
  float* s_sig = s_tgt + kMcMaxG * kPad;                     // [kMcMaxQ][kPad]  sigmoid(x[q][c]), row Q = 1
  const int t = threadIdx.x;
  for (int i = t; i < (kMcMaxQ + kMcK + kMcMaxG + kMcMaxQ) * kPad; i += kThreads) s_coeff[i] = 0.f;
  __syncthreads();
  for (int i = t; i < Q * K; i += kThreads) s_coeff[(i / K) * kPad + (i % K)] = coeff[i];

  const int64_t c_begin = blockIdx.x * cols_per_cta;
  const int64_t c_end = min(N, c_begin + cols_per_cta);
  const int rows = Q + 1;                                     // incl. the all-ones row
  const int k_end = (K + 3) & ~3;
  float neg_sum[2][4] = {};
  float acc_s[4][4] = {};
  float acc_pt[2] = {0.f, 0.f};
  const int q4_2 = t >> 2, g4_2 = t & 3;                     // stage-2 tile of this thread
  const int k_3 = t >> 3, g2_3 = t & 7;                      // stage-3 outputs of this thread: PT[k_3][2 g2_3], PT[k_3][2 g2_3 + 1]

  for (int64_t c0 = c_begin; c0 < c_end; c0 += kTC) {
    __syncthreads();                                          // previous iteration's stages 2 / 3 are done with the tiles
    {
      const int c = t & 31;
      const bool ok = c0 + c < c_end;
      for (int k = t >> 5; k < kMcK; k += 8) s_proto[k * kPad + c] = (ok && k < K) ? __ldg(proto + k * N + c0 + c) : 0.f;
      for (int g = t >> 5; g < kMcMaxG; g += 8)
        s_tgt[g * kPad + c] = !ok ? 0.f : (g < G ? __ldg(tgt + g * N + c0 + c) : (g == G ? 1.f : 0.f));
    }
    __syncthreads();
    // stage 1: x = coeff . proto for a 4 x 4 (query, column) tile, then sigmoid / softplus
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int item = t + kThreads * pass;
      const int q4 = item >> 3, c4 = item & 7;
      if (q4 * 4 < rows) {
        float x[4][4] = {};
        for (int k = 0; k < k_end; k += 4) {
          float4 a[4], b[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = lds4(s_coeff + (q4 * 4 + r) * kPad + k);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) b[kk] = lds4(s_proto + (k + kk) * kPad + c4 * 4);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float av[4] = {a[r].x, a[r].y, a[r].z, a[r].w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              x[r][0] = fmaf(av[kk], b[kk].x, x[r][0]);
              x[r][1] = fmaf(av[kk], b[kk].y, x[r][1]);
              x[r][2] = fmaf(av[kk], b[kk].z, x[r][2]);
              x[r][3] = fmaf(av[kk], b[kk].w, x[r][3]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int q = q4 * 4 + r;
          float sg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = c0 + c4 * 4 + j < c_end;
            if (q < Q && ok) {
              const float v = x[r][j];
              // binary_cross_entropy_with_logits(x, 0) = max(x, 0) + log1p(exp(-|x|));  sigmoid from the same exponential.
              // Hardware exp2 / log2 / reciprocal (2 ulp; log(1 + e) is off by <= 6e-8 absolute for tiny e): the costs are
              // sums over >= 10^4 such terms of magnitude ~0.5, far inside the 1e-4 bar (tests/test_consumers_gpu.py)
              const float e = __expf(-fabsf(v));
              const float r1 = __frcp_rn(1.f + e);
              neg_sum[pass][r] += fmaxf(v, 0.f) + __logf(1.f + e);
              sg[j] = v >= 0.f ? r1 : e * r1;
            } else {
              sg[j] = (q == Q && ok) ? 1.f : 0.f;
            }
          }
          if (q < kMcMaxQ) *reinterpret_cast<float4*>(s_sig + q * kPad + c4 * 4) = make_float4(sg[0], sg[1], sg[2], sg[3]);
        }
      }
    }
    // stage 3 (reads only s_proto / s_tgt, so it can run before the barrier): PT[k][g] += proto[k][:] . tgt[g][:]
    if (g2_3 * 2 < G) {
#pragma unroll
      for (int c4 = 0; c4 < kTC / 4; ++c4) {
        const float4 p = lds4(s_proto + k_3 * kPad + c4 * 4);
        const float4 t0 = lds4(s_tgt + (g2_3 * 2) * kPad + c4 * 4), t1 = lds4(s_tgt + (g2_3 * 2 + 1) * kPad + c4 * 4);
        acc_pt[0] += p.x * t0.x + p.y * t0.y + p.z * t0.z + p.w * t0.w;
        acc_pt[1] += p.x * t1.x + p.y * t1.y + p.z * t1.z + p.w * t1.w;
      }
    }
    __syncthreads();
    // stage 2: sigmoid (rows) x targets (columns), 4 x 4 tile per thread
    if (q4_2 * 4 < rows && g4_2 * 4 <= G) {
#pragma unroll
      for (int c4 = 0; c4 < kTC / 4; ++c4) {
        float4 sg[4], tg[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          sg[r] = lds4(s_sig + (q4_2 * 4 + r) * kPad + c4 * 4);
          tg[r] = lds4(s_tgt + (g4_2 * 4 + r) * kPad + c4 * 4);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int g = 0; g < 4; ++g)
            acc_s[r][g] += sg[r].x * tg[g].x + sg[r].y * tg[g].y + sg[r].z * tg[g].z + sg[r].w * tg[g].w;
      }
    }
  }
  // partial sums of this CTA -> workspace
  if (q4_2 * 4 < rows && g4_2 * 4 <= G) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int q = q4_2 * 4 + r, gg = g4_2 * 4 + g;
        if (q < rows && gg <= G) atomicAdd(ws + q * kMcMaxG + gg, acc_s[r][g]);
      }
  }
  if (g2_3 * 2 < G && k_3 < K) {
    atomicAdd(ws + kMcMaxQ * kMcMaxG + k_3 * kMcMaxG + g2_3 * 2, acc_pt[0]);
    if (g2_3 * 2 + 1 < G) atomicAdd(ws + kMcMaxQ * kMcMaxG + k_3 * kMcMaxG + g2_3 * 2 + 1, acc_pt[1]);
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int q4 = (t + kThreads * pass) >> 3;
    // the 8 threads that share a query group hold partial sums over different columns: fold them first (consecutive lanes)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v = neg_sum[pass][r];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      const int q = q4 * 4 + r;
      if ((t & 7) == 0 && q < Q) atomicAdd(ws + kMcMaxQ * kMcMaxG + kMcK * kMcMaxG + q, v);
    }
  }
}

__global__ void match_cost_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ coeff, int Q, int K, int G, int64_t N, int ld,
                                           float* __restrict__ cost_bce, float* __restrict__ cost_dice) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q * G) return;
  const int q = i / G, g = i % G;
  const float* pt = ws + kMcMaxQ * kMcMaxG;
  float xt = 0.f;                                               // sum_c x[q,c] tgt[g,c] = sum_k coeff[q,k] PT[k,g]
  for (int k = 0; k < K; ++k) xt = fmaf(coeff[q * K + k], pt[k * kMcMaxG + g], xt);
  const float st = ws[q * kMcMaxG + g];
  const float sig_sum = ws[q * kMcMaxG + G], tgt_sum = ws[Q * kMcMaxG + g], neg = ws[kMcMaxQ * kMcMaxG + kMcK * kMcMaxG + q];
  // pos * t + neg * (1 - t) = neg - x * t  (pos - neg = -x), matcher.py:58-61
  cost_bce[q * ld + g] = (neg - xt) / static_cast<float>(N);
  cost_dice[q * ld + g] = 1.f - (2.f * st + 1.f) / (sig_sum + tgt_sum + 1.f);           // matcher.py:25-27
}

// ------------------------------------------------------------------------------------------ NMS soft IoU
constexpr int kSiMaxQ = 128;                                   // rows per chunk incl. the all-ones row
constexpr size_t kSiSmem = static_cast<size_t>(2 * kSiMaxQ) * kPad * sizeof(float);

__global__ void __launch_bounds__(kThreads)
nms_siou_kernel(const float* __restrict__ mask, int i0, int Qi, int j0, int Qj, int T, int H, int W, int t_step, int T2, int H2, int W2,
                int64_t cols_per_cta, float* __restrict__ ws) {
  extern __shared__ float4 smem4[];
  float* s_soft = reinterpret_cast<float*>(smem4);            // [kSiMaxQ][kPad] sigmoid of rows i0.., row Qi = 1
  float* s_hard = s_soft + kSiMaxQ * kPad;                    // [kSiMaxQ][kPad] (logit > 0) of rows j0.., row Qj = 1
  const int t = threadIdx.x;
  for (int i = t; i < 2 * kSiMaxQ * kPad; i += kThreads) s_soft[i] = 0.f;
  const int64_t N2 = static_cast<int64_t>(T2) * H2 * W2;
  const int64_t c_begin = blockIdx.x * cols_per_cta, c_end = min(N2, c_begin + cols_per_cta);
  const int64_t plane = static_cast<int64_t>(H) * W;
  float acc[4][4][4] = {};
  for (int64_t c0 = c_begin; c0 < c_end; c0 += kTC) {
    __syncthreads();
    {
      const int c = t & 31;
      const int64_t cc = c0 + c;
      const bool ok = cc < c_end;
      int64_t src = 0;
      if (ok) {                                               // nearest, scale 0.5: source pixel (2y, 2x); frames t_step apart
        const int x2 = static_cast<int>(cc % W2), y2 = static_cast<int>((cc / W2) % H2), t2 = static_cast<int>(cc / (static_cast<int64_t>(W2) * H2));
        src = static_cast<int64_t>(t2) * t_step * plane + static_cast<int64_t>(2 * y2) * W + 2 * x2;
      }
      const int r_end = max(Qi, Qj) + 1;                      // rows beyond stay zero from the initial fill
      for (int r = t >> 5; r < r_end; r += 8) {
        float so = 0.f, ha = 0.f;
        if (ok) {
          if (r < Qi) so = 1.f / (1.f + expf(-__ldg(mask + static_cast<int64_t>(i0 + r) * T * plane + src)));
          else if (r == Qi) so = 1.f;
          // mask_soft.gt(0.5) on the fp32 sigmoid (mdqe.py:388-389), not logit > 0: they differ for tiny positive logits
          if (r < Qj) ha = (1.f / (1.f + expf(-__ldg(mask + static_cast<int64_t>(j0 + r) * T * plane + src)))) > 0.5f ? 1.f : 0.f;
          else if (r == Qj) ha = 1.f;
        }
        s_soft[r * kPad + c] = so;
        s_hard[r * kPad + c] = ha;
      }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int item = t + kThreads * p;
      const int i4 = item >> 5, j4 = item & 31;
      if (i4 * 4 <= Qi && j4 * 4 <= Qj) {
#pragma unroll
        for (int c4 = 0; c4 < kTC / 4; ++c4) {
          float4 so[4], ha[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            so[r] = lds4(s_soft + (i4 * 4 + r) * kPad + c4 * 4);
            ha[r] = lds4(s_hard + (j4 * 4 + r) * kPad + c4 * 4);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int g = 0; g < 4; ++g)
              acc[p][r][g] += so[r].x * ha[g].x + so[r].y * ha[g].y + so[r].z * ha[g].z + so[r].w * ha[g].w;
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int item = t + kThreads * p;
    const int i4 = item >> 5, j4 = item & 31;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int i = i4 * 4 + r, j = j4 * 4 + g;
        if (i <= Qi && j <= Qj) atomicAdd(ws + i * kSiMaxQ + j, acc[p][r][g]);
      }
  }
}

__global__ void nms_siou_finalize_kernel(const float* __restrict__ ws, int i0, int Qi, int j0, int Qj, int Q, float* __restrict__ siou) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Qi * Qj) return;
  const int i = idx / Qj, j = idx % Qj;
  const float num = ws[i * kSiMaxQ + j];
  const float den = ws[i * kSiMaxQ + Qj] + ws[Qi * kSiMaxQ + j] - num;      // mdqe.py:393
  siou[static_cast<int64_t>(i0 + i) * Q + j0 + j] = num / (den + 1.f);
}

// ------------------------------------------------------------------------------------------ aligned_bilinear (+ sigmoid)
// out[y, x] = bilinear(in, max(y - f/2, 0) / f, max(x - f/2, 0) / f) with the source index clamped at the last row / column:
// replicate-pad by one, align_corners=True resize to (f*h + 1, f*w + 1), replicate-pad f/2 at the top / left, crop (misc.py:494-507).
// block = 64 quads of 4 output pixels (x) by 16 output rows of ONE image (4 consecutive rows per thread: the column geometry
// is computed once per thread, consecutive rows mostly share their two source rows); 32-bit index arithmetic
constexpr int kAbRows = 16;
__global__ void __launch_bounds__(kThreads)
aligned_bilinear_kernel(const float* __restrict__ in, int blocks_per_img, int H, int W, int f, int do_sigmoid, float* __restrict__ out) {
  const int OW = W * f, OH = H * f;
  const int xq = blockIdx.y * 64 + threadIdx.x;
  if (xq * 4 >= OW) return;
  const unsigned img = blockIdx.x / static_cast<unsigned>(blocks_per_img);
  const int y_base = (blockIdx.x - img * blocks_per_img) * kAbRows + threadIdx.y * 4;
  const float inv_f = 1.f / static_cast<float>(f);
  int xa[4], xb[4];
  float lx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = xq * 4 + j;
    const float sx = static_cast<float>(max(x - f / 2, 0)) * inv_f;
    const int x0 = static_cast<int>(sx);
    lx[j] = sx - static_cast<float>(x0);
    xa[j] = min(x0, W - 1);
    xb[j] = min(x0 + 1, W - 1);
  }
  const float* src = in + static_cast<size_t>(img) * H * W;
  float* dst_img = out + static_cast<size_t>(img) * OH * OW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = y_base + k;
    if (y >= OH) return;
    const float sy = static_cast<float>(max(y - f / 2, 0)) * inv_f;
    const int y0 = static_cast<int>(sy);
    const float ly = sy - static_cast<float>(y0);
    const float* r0 = src + min(y0, H - 1) * W;
    const float* r1 = src + min(y0 + 1, H - 1) * W;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // same operation order as upsample_bilinear2d: rows first, then the two rows blended
      const float top = (1.f - lx[j]) * __ldg(r0 + xa[j]) + lx[j] * __ldg(r0 + xb[j]);
      const float bot = (1.f - lx[j]) * __ldg(r1 + xa[j]) + lx[j] * __ldg(r1 + xb[j]);
      float o = (1.f - ly) * top + ly * bot;
      if (do_sigmoid) o = __frcp_rn(1.f + expf(-o));          // correctly rounded reciprocal == 1.f / x
      v[j] = o;
    }
    float* dst = dst_img + static_cast<size_t>(y) * OW + xq * 4;
    if ((OW & 3) == 0) __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));    // streaming: written once, 16x the input
    else
      for (int j = 0; j < 4 && xq * 4 + j < OW; ++j) dst[j] = v[j];
  }
}

// ------------------------------------------------------------------------------------------ query initialisation sampling
struct BorderSample {
  int x0, y0, x1, y1;
  float lx, ly, gx, gy;          // gx / gy: d(pixel coordinate) / d(normalised coordinate), 0 where the border clamp is active
};
// grid_sample(bilinear, padding_mode="border", align_corners=False) for a normalised coordinate c in [0, 1] (grid = 2c - 1):
// pixel = c * size - 0.5, clipped to [0, size - 1] (GridSampler.h clip_coordinates), corners floor / floor + 1.
__device__ __forceinline__ BorderSample border_sample(float cx, float cy, int H, int W) {
  BorderSample s;
  float x = cx * static_cast<float>(W) - 0.5f, y = cy * static_cast<float>(H) - 0.5f;
  s.gx = (x > 0.f && x < static_cast<float>(W - 1)) ? static_cast<float>(W) : 0.f;
  s.gy = (y > 0.f && y < static_cast<float>(H - 1)) ? static_cast<float>(H) : 0.f;
  x = fminf(static_cast<float>(W - 1), fmaxf(x, 0.f));
  y = fminf(static_cast<float>(H - 1), fmaxf(y, 0.f));
  const float fx = floorf(x), fy = floorf(y);
  s.x0 = static_cast<int>(fx); s.y0 = static_cast<int>(fy);
  s.x1 = s.x0 + 1; s.y1 = s.y0 + 1;
  s.lx = x - fx; s.ly = y - fy;
  return s;
}

// one warp per (b, q); lanes walk the channels in float4 steps
template <bool BWD>
__global__ void __launch_bounds__(kThreads)
query_init_kernel(const float* __restrict__ feat, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                  const float* __restrict__ coords, const float* __restrict__ grad_out, int B, int S, int C, int L, int Q,
                  float* __restrict__ out, float* __restrict__ grad_feat, float* __restrict__ grad_coords) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * Q) return;
  const int b = warp / Q;
  const float cx = __ldg(coords + 2 * warp), cy = __ldg(coords + 2 * warp + 1);
  const float inv_l = 1.f / static_cast<float>(L);
  float gcx = 0.f, gcy = 0.f;
  for (int cb = lane * 4; cb < C; cb += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), go = acc;
    if (BWD) go = __ldg(reinterpret_cast<const float4*>(grad_out + static_cast<int64_t>(warp) * C + cb));
    for (int l = 0; l < L; ++l) {
      const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
      const BorderSample s = border_sample(cx, cy, H, W);
      const int64_t base = (static_cast<int64_t>(b) * S + lsi[l]) * C + cb;
      const bool okx = s.x1 < W, oky = s.y1 < H;              // the far corner leaves the map only when its weight is 0
      const int64_t o00 = base + (static_cast<int64_t>(s.y0) * W + s.x0) * C;
      const int64_t o01 = o00 + C, o10 = o00 + static_cast<int64_t>(W) * C, o11 = o10 + C;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(feat + o00));
      const float4 v01 = okx ? __ldg(reinterpret_cast<const float4*>(feat + o01)) : z;
      const float4 v10 = oky ? __ldg(reinterpret_cast<const float4*>(feat + o10)) : z;
      const float4 v11 = (okx && oky) ? __ldg(reinterpret_cast<const float4*>(feat + o11)) : z;
      const float w00 = (1.f - s.lx) * (1.f - s.ly), w01 = s.lx * (1.f - s.ly), w10 = (1.f - s.lx) * s.ly, w11 = s.lx * s.ly;
      if (!BWD) {
        acc.x += w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
        acc.y += w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
        acc.z += w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
        acc.w += w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
      } else {
        const float g[4] = {go.x * inv_l, go.y * inv_l, go.z * inv_l, go.w * inv_l};
        const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
        const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
        float dx = 0.f, dy = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          atomicAdd(grad_feat + o00 + j, w00 * g[j]);
          if (okx) atomicAdd(grad_feat + o01 + j, w01 * g[j]);
          if (oky) atomicAdd(grad_feat + o10 + j, w10 * g[j]);
          if (okx && oky) atomicAdd(grad_feat + o11 + j, w11 * g[j]);
          dx += g[j] * ((a01[j] - a00[j]) * (1.f - s.ly) + (a11[j] - a10[j]) * s.ly);
          dy += g[j] * ((a10[j] - a00[j]) * (1.f - s.lx) + (a11[j] - a01[j]) * s.lx);
        }
        gcx += dx * s.gx;
        gcy += dy * s.gy;
      }
    }
    if (!BWD)
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(warp) * C + cb) = make_float4(acc.x * inv_l, acc.y * inv_l, acc.z * inv_l, acc.w * inv_l);
  }
  if (BWD) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      gcx += __shfl_xor_sync(0xffffffffu, gcx, o);
      gcy += __shfl_xor_sync(0xffffffffu, gcy, o);
    }
    if (lane == 0) { grad_coords[2 * warp] = gcx; grad_coords[2 * warp + 1] = gcy; }
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

size_t mask_match_cost_workspace_bytes(void) { return kMcWsFloats * sizeof(float); }

int mask_match_cost(void* stream, const void* coeff, const void* proto, const void* targets, int Q, int K, int G, int64_t Ncols,
                    void* workspace, void* cost_bce, void* cost_dice) {
  if (Q < 0 || G < 0 || K <= 0 || K > 32 || Ncols <= 0) return fail(MSDA_ERR_INVALID_ARG, "mask_match_cost: Q=%d K=%d G=%d Ncols=%lld (K <= 32)", Q, K, G, (long long)Ncols);
  if (Q == 0 || G == 0) return 0;
  if (!coeff || !proto || !targets || !workspace || !cost_bce || !cost_dice) return fail(MSDA_ERR_INVALID_ARG, "mask_match_cost: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr = false;
  if (!attr) {
    if (int rc = check_cuda(cudaFuncSetAttribute(match_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMcSmem)), "cudaFuncSetAttribute")) return rc;
    attr = true;
  }
  const int64_t chunks = (Ncols + kTC - 1) / kTC;
  const int slots = 3 * sm_count();                           // three resident CTAs per SM
  const int ctas = static_cast<int>(chunks < slots ? chunks : slots);
  const int64_t cols_per_cta = ((chunks + ctas - 1) / ctas) * kTC;
  for (int q0 = 0; q0 < Q; q0 += kMcMaxQ - 1) {
    const int qn = Q - q0 < kMcMaxQ - 1 ? Q - q0 : kMcMaxQ - 1;
    for (int g0 = 0; g0 < G; g0 += kMcMaxG - 1) {
      const int gn = G - g0 < kMcMaxG - 1 ? G - g0 : kMcMaxG - 1;
      if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, kMcWsFloats * sizeof(float), st), "cudaMemsetAsync(workspace)")) return rc;
      {
        ProfScope prof(st, 4, static_cast<int64_t>(qn) * Ncols);
        match_cost_kernel<<<ctas, kThreads, kMcSmem, st>>>(static_cast<const float*>(coeff) + static_cast<int64_t>(q0) * K, static_cast<const float*>(proto),
                                                           static_cast<const float*>(targets) + static_cast<int64_t>(g0) * Ncols, qn, K, gn, Ncols,
                                                           cols_per_cta, static_cast<float*>(workspace));
      }
      if (int rc = after_launch("match_cost_kernel")) return rc;
      match_cost_finalize_kernel<<<(qn * gn + 255) / 256, 256, 0, st>>>(static_cast<const float*>(workspace),
                                                                         static_cast<const float*>(coeff) + static_cast<int64_t>(q0) * K, qn, K, gn, Ncols, G,
                                                                         static_cast<float*>(cost_bce) + static_cast<int64_t>(q0) * G + g0,
                                                                         static_cast<float*>(cost_dice) + static_cast<int64_t>(q0) * G + g0);
      if (int rc = after_launch("match_cost_finalize_kernel")) return rc;
    }
  }
  return 0;
}

size_t mask_nms_siou_workspace_bytes(void) { return static_cast<size_t>(kSiMaxQ) * kSiMaxQ * sizeof(float); }

int mask_nms_siou(void* stream, const void* mask_pred, int Q, int T, int H, int W, void* workspace, void* siou) {
  if (Q < 0 || T <= 0 || H < 2 || W < 2) return fail(MSDA_ERR_INVALID_ARG, "mask_nms_siou: Q=%d T=%d H=%d W=%d", Q, T, H, W);
  if (Q == 0) return 0;
  if (!mask_pred || !workspace || !siou) return fail(MSDA_ERR_INVALID_ARG, "mask_nms_siou: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr = false;
  if (!attr) {
    if (int rc = check_cuda(cudaFuncSetAttribute(nms_siou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSiSmem)), "cudaFuncSetAttribute")) return rc;
    attr = true;
  }
  const int t_step = T >= 5 ? 2 : 1;                            // mask_pred[:, ::2] if T >= 5 (mdqe.py:386)
  const int T2 = (T + t_step - 1) / t_step, H2 = H / 2, W2 = W / 2;
  const int64_t N2 = static_cast<int64_t>(T2) * H2 * W2;
  const int64_t chunks = (N2 + kTC - 1) / kTC;
  const int ctas = static_cast<int>(chunks < sm_count() ? chunks : sm_count());
  const int64_t cols_per_cta = ((chunks + ctas - 1) / ctas) * kTC;
  for (int i0 = 0; i0 < Q; i0 += kSiMaxQ - 1) {
    const int qi = Q - i0 < kSiMaxQ - 1 ? Q - i0 : kSiMaxQ - 1;
    for (int j0 = 0; j0 < Q; j0 += kSiMaxQ - 1) {
      const int qj = Q - j0 < kSiMaxQ - 1 ? Q - j0 : kSiMaxQ - 1;
      if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, mask_nms_siou_workspace_bytes(), st), "cudaMemsetAsync(workspace)")) return rc;
      {
        ProfScope prof(st, 5, static_cast<int64_t>(qi) * N2);
        nms_siou_kernel<<<ctas, kThreads, kSiSmem, st>>>(static_cast<const float*>(mask_pred), i0, qi, j0, qj, T, H, W, t_step, T2, H2, W2, cols_per_cta,
                                                         static_cast<float*>(workspace));
      }
      if (int rc = after_launch("nms_siou_kernel")) return rc;
      nms_siou_finalize_kernel<<<(qi * qj + 255) / 256, 256, 0, st>>>(static_cast<const float*>(workspace), i0, qi, j0, qj, Q, static_cast<float*>(siou));
      if (int rc = after_launch("nms_siou_finalize_kernel")) return rc;
    }
  }
  return 0;
}

int aligned_bilinear_sigmoid(void* stream, const void* in, int64_t n_img, int H, int W, int factor, int apply_sigmoid, void* out) {
  if (n_img < 0 || H <= 0 || W <= 0 || factor < 1) return fail(MSDA_ERR_INVALID_ARG, "aligned_bilinear_sigmoid: n=%lld H=%d W=%d factor=%d", (long long)n_img, H, W, factor);
  if (n_img == 0) return 0;
  if (!in || !out) return fail(MSDA_ERR_INVALID_ARG, "aligned_bilinear_sigmoid: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks_per_img = (H * factor + kAbRows - 1) / kAbRows;
  const int quads = (W * factor + 3) / 4;
  if (n_img * blocks_per_img > 0x7fffffff) return fail(MSDA_ERR_INVALID_ARG, "aligned_bilinear_sigmoid: too many rows");
  {
    ProfScope prof(st, 6, n_img * H * factor * quads);
    aligned_bilinear_kernel<<<dim3(static_cast<unsigned>(n_img * blocks_per_img), (quads + 63) / 64), dim3(64, 4), 0, st>>>(
        static_cast<const float*>(in), blocks_per_img, H, W, factor, apply_sigmoid, static_cast<float*>(out));
  }
  return after_launch("aligned_bilinear_kernel");
}

static int query_init_check(const char* who, const void* feat, const int64_t* shapes, const int64_t* lsi, const void* coords, int B, int S, int C, int L, int Q) {
  if (B < 0 || S <= 0 || C <= 0 || (C & 3) || L <= 0 || Q < 0) return fail(MSDA_ERR_INVALID_ARG, "%s: B=%d S=%d C=%d L=%d Q=%d (C must be a multiple of 4)", who, B, S, C, L, Q);
  if (B * Q > 0 && (!feat || !shapes || !lsi || !coords)) return fail(MSDA_ERR_INVALID_ARG, "%s: NULL pointer", who);
  return 0;
}

int query_init_sample_forward(void* stream, const void* feat, const int64_t* shapes, const int64_t* level_start, const void* coords, int B, int S,
                              int C, int L, int Q, void* out) {
  if (int rc = query_init_check("query_init_sample_forward", feat, shapes, level_start, coords, B, S, C, L, Q)) return rc;
  if (B * Q == 0) return 0;
  if (!out) return fail(MSDA_ERR_INVALID_ARG, "query_init_sample_forward: out is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps = B * Q;
  query_init_kernel<false><<<(warps + 7) / 8, kThreads, 0, st>>>(static_cast<const float*>(feat), shapes, level_start, static_cast<const float*>(coords),
                                                                 nullptr, B, S, C, L, Q, static_cast<float*>(out), nullptr, nullptr);
  return after_launch("query_init_kernel<fwd>");
}

int query_init_sample_backward(void* stream, const void* feat, const int64_t* shapes, const int64_t* level_start, const void* coords,
                               const void* grad_out, int B, int S, int C, int L, int Q, void* grad_feat, void* grad_coords) {
  if (int rc = query_init_check("query_init_sample_backward", feat, shapes, level_start, coords, B, S, C, L, Q)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_feat && B > 0)
    if (int rc = check_cuda(cudaMemsetAsync(grad_feat, 0, static_cast<size_t>(B) * S * C * sizeof(float), st), "cudaMemsetAsync(grad_feat)")) return rc;
  if (B * Q == 0) return 0;
  if (!grad_out || !grad_feat || !grad_coords) return fail(MSDA_ERR_INVALID_ARG, "query_init_sample_backward: NULL pointer");
  const int warps = B * Q;
  query_init_kernel<true><<<(warps + 7) / 8, kThreads, 0, st>>>(static_cast<const float*>(feat), shapes, level_start, static_cast<const float*>(coords),
                                                                static_cast<const float*>(grad_out), B, S, C, L, Q, nullptr, static_cast<float*>(grad_feat),
                                                                static_cast<float*>(grad_coords));
  return after_launch("query_init_kernel<bwd>");
}

}  // extern "C"
