// Callers either side of the hot path, sm_100a SIMT kernels behind the C ABI of include/msda_b200.h.
// Paths relative to /root/reference:
//   mask_match_cost          mdqe/models/matcher.py:182-200 with batch_sigmoid_ce_loss (:36-61) and batch_dice_loss (:11-28):
//                            the Hungarian matcher's mask costs, fused with the mask contraction -- out_masks [Q, T*H*W]
//                            (48 MB per clip at R50_ovis_360) is never written, proto and the targets are read once;
//   mask_losses_*            mdqe/models/criterion.py:440-473: BCE + dice (plain or inter-instance) of the matched rows, forward and
//                            backward, fused with the contraction of those rows;
//   mask_nms_siou            mdqe/mdqe.py:394-407: soft-IoU matrix of inference_clip (frame stride 2 for clips of 5+ frames,
//                            nearest 0.5x downsampling, sigmoid, threshold, Q x Q product) in one pass over mask_pred;
//   mask_track_siou          mdqe/tracking/OverTracker.py:92-113: hard-mask IoU between the tracker's memory and a new clip;
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

// ------------------------------------------------------------------------------------------ matched-mask losses
// criterion.py:440-473 with sigmoid_ce_loss / dice_loss (:20-43, :87-108) or the inter-instance forms (:51-81, :116-145), fused with the
// contraction of the G matched coefficient rows: x[g,c] = coeff[g,:] . proto[:,c] is recomputed per 32 plane columns in both passes
// and never written.  Per row g the forward needs 7 sums over the plane (kept in row_stats for the backward):
//   0: sum l*w   (l = softplus(x) - x*t, w = ti + 1)      1: sum w      2: sum s*t      3: sum (1 - s)*tib      4: sum s      5: sum t
//   6: sum tib   (tib = ti > 0.5 and 1 - t > 0.5)
constexpr int kMlMaxG = 32, kMlStats = 8;

__device__ __forceinline__ void ml_point(float x, float t, float ti, float& s, float& l, float& w, float& tib) {
  const float e = expf(-fabsf(x));
  const float r = 1.f / (1.f + e);
  s = x >= 0.f ? r : e * r;
  l = fmaxf(x, 0.f) + log1pf(e) - x * t;                       // binary_cross_entropy_with_logits(x, t)
  w = ti + 1.f;                                                // criterion.py:140
  tib = (ti > 0.5f && (1.f - t) > 0.5f) ? 1.f : 0.f;           // :69
}

// thread (c = t % 32, rows g = t / 32 + 8 j): the proto column lives in registers and is reused by the thread's 4 rows
__global__ void __launch_bounds__(kThreads)
mask_losses_fwd_kernel(const float* __restrict__ coeff, const float* __restrict__ proto, const float* __restrict__ tgt,
                       const float* __restrict__ tgt_inter, int G, int K, int64_t N, int64_t cols_per_cta, float* __restrict__ ws) {
  __shared__ float s_coeff[kMlMaxG][33];
  const int t = threadIdx.x, c = t & 31, g0 = t >> 5;
  for (int i = t; i < kMlMaxG * 32; i += kThreads) s_coeff[i >> 5][i & 31] = ((i >> 5) < G && (i & 31) < K) ? coeff[(i >> 5) * K + (i & 31)] : 0.f;
  __syncthreads();
  const int64_t c_begin = blockIdx.x * cols_per_cta, c_end = min(N, c_begin + cols_per_cta);
  float acc[4][7] = {};
  for (int64_t c0 = c_begin; c0 < c_end; c0 += kTC) {
    const int64_t cc = c0 + c;
    if (cc >= c_end) continue;
    float p[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) p[k] = k < K ? __ldg(proto + k * N + cc) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = g0 + 8 * j;
      if (g < G) {
        float x = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) x = fmaf(s_coeff[g][k], p[k], x);
        const float tv = __ldg(tgt + g * N + cc), ti = tgt_inter ? __ldg(tgt_inter + g * N + cc) : 0.f;
        float sg, l, w, tib;
        ml_point(x, tv, ti, sg, l, w, tib);
        acc[j][0] += l * w; acc[j][1] += w; acc[j][2] += sg * tv; acc[j][3] += (1.f - sg) * tib; acc[j][4] += sg; acc[j][5] += tv; acc[j][6] += tib;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      float v = acc[j][i];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (c == 0 && g0 + 8 * j < G) atomicAdd(ws + (g0 + 8 * j) * kMlStats + i, v);
    }
}

__global__ void mask_losses_finalize_kernel(const float* __restrict__ ws, int G, int64_t N, int has_inter, float num_masks, float* __restrict__ row_stats,
                                            float* __restrict__ losses) {
  __shared__ float s_b[kMlMaxG], s_d[kMlMaxG];
  const int g = threadIdx.x;
  if (g < G) {
    const float* r = ws + g * kMlStats;
    // plain forms: loss.mean(1) (:108) and no tib terms -- the same expressions with w == 1 (sum w = N) and tib == 0
    const float wsum = has_inter ? fmaxf(r[1], 1.f) : static_cast<float>(N);
    s_b[g] = r[0] / wsum;                                                         // :142 / :108
    const float num = 2.f * r[2] + r[3], den = r[4] + r[5] + r[6];                // :77-78 / :39-40
    s_d[g] = 1.f - (num + 1.f) / (den + 1.f);                                     // :79 / :41
    for (int i = 0; i < kMlStats; ++i) row_stats[g * kMlStats + i] = i < 7 ? r[i] : 0.f;
  }
  __syncthreads();
  if (g == 0) {
    float b = 0.f, d = 0.f;
    for (int i = 0; i < G; ++i) { b += s_b[i]; d += s_d[i]; }
    const float inv = 1.f / fmaxf(num_masks, 1.f);                                // :144 / :81
    losses[0] = b * inv;
    losses[1] = d * inv;
  }
}

// backward: gx[g,c] = d(g_mask * loss_mask + g_dice * loss_dice) / d x[g,c], then grad_proto = coeff^T . gx (written per column tile)
// and grad_coeff = gx . proto^T (accumulated per CTA, atomics at the end)
__global__ void __launch_bounds__(kThreads)
mask_losses_bwd_kernel(const float* __restrict__ coeff, const float* __restrict__ proto, const float* __restrict__ tgt,
                       const float* __restrict__ tgt_inter, const float* __restrict__ row_stats, const float* __restrict__ grad_losses,
                       int G, int K, int64_t N, int64_t cols_per_cta, float num_masks, float* __restrict__ grad_coeff_ws, float* __restrict__ grad_proto) {
  __shared__ float s_coeff[kMlMaxG][33];
  __shared__ __align__(16) float s_gx[kMlMaxG][kPad];
  __shared__ __align__(16) float s_proto[32][kPad];
  __shared__ float s_cb[kMlMaxG], s_cd1[kMlMaxG], s_cd2[kMlMaxG];      // per-row constants of the two gradients
  const int t = threadIdx.x, c = t & 31, g0 = t >> 5;
  for (int i = t; i < kMlMaxG * 32; i += kThreads) s_coeff[i >> 5][i & 31] = ((i >> 5) < G && (i & 31) < K) ? coeff[(i >> 5) * K + (i & 31)] : 0.f;
  for (int i = t; i < kMlMaxG * kPad; i += kThreads) (&s_gx[0][0])[i] = 0.f;
  if (t < kMlMaxG) {
    float cb = 0.f, cd1 = 0.f, cd2 = 0.f;
    if (t < G) {
      const float* r = row_stats + t * kMlStats;
      const float inv = 1.f / fmaxf(num_masks, 1.f);
      const float wsum = tgt_inter ? fmaxf(r[1], 1.f) : static_cast<float>(N);
      const float num = 2.f * r[2] + r[3], den = r[4] + r[5] + r[6];
      cb = grad_losses[0] * inv / wsum;                                 // d loss_mask / d x = cb * (s - t) * w
      // d loss_dice / d x = -s(1-s) [ (2t - tib)(den + 1) - (num + 1) ] / (den + 1)^2 = s(1-s) [ cd2 - cd1 (2t - tib) ]
      cd1 = grad_losses[1] * inv / (den + 1.f);
      cd2 = grad_losses[1] * inv * (num + 1.f) / ((den + 1.f) * (den + 1.f));
    }
    s_cb[t] = cb; s_cd1[t] = cd1; s_cd2[t] = cd2;
  }
  __syncthreads();
  const int64_t c_begin = blockIdx.x * cols_per_cta, c_end = min(N, c_begin + cols_per_cta);
  const int gB = t >> 3, k4B = t & 7;                                   // grad_coeff tile of this thread: row gB, k = 4 k4B .. +3
  float acc_gc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t c0 = c_begin; c0 < c_end; c0 += kTC) {
    const int64_t cc = c0 + c;
    const bool ok = cc < c_end;
    __syncthreads();                                                    // previous iteration is done with s_gx / s_proto
    float p[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      p[k] = (ok && k < K) ? __ldg(proto + k * N + cc) : 0.f;
      if (g0 == (k & 7)) s_proto[k][c] = p[k];                          // each of the 8 warps stores 4 of the 32 rows
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = g0 + 8 * j;
      float gx = 0.f;
      if (g < G && ok) {
        float x = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) x = fmaf(s_coeff[g][k], p[k], x);
        const float tv = __ldg(tgt + g * N + cc), ti = tgt_inter ? __ldg(tgt_inter + g * N + cc) : 0.f;
        float sg, l, w, tib;
        ml_point(x, tv, ti, sg, l, w, tib);
        gx = s_cb[g] * (sg - tv) * w + sg * (1.f - sg) * (s_cd2[g] - s_cd1[g] * (2.f * tv - tib));
      }
      s_gx[g][c] = gx;
    }
    __syncthreads();
    {                                                                   // grad_proto[k][c0 + 4 c4 ..] = sum_g coeff[g][k] gx[g][..]
      const int k = t >> 3, c4 = t & 7;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < G; ++g) {
        const float a = s_coeff[g][k];
        const float4 v = lds4(&s_gx[g][c4 * 4]);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
      if (k < K) {
        const int64_t col = c0 + c4 * 4;
        float* dst = grad_proto + k * N + col;
        if (col + 3 < c_end && (N & 3) == 0) *reinterpret_cast<float4*>(dst) = o;
        else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
          for (int i = 0; i < 4 && col + i < c_end; ++i) dst[i] = ov[i];
        }
      }
    }
    if (gB < G) {                                                       // grad_coeff[g][k] += sum_c gx[g][c] proto[k][c]
#pragma unroll
      for (int c4 = 0; c4 < kTC / 4; ++c4) {
        const float4 v = lds4(&s_gx[gB][c4 * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 q = lds4(&s_proto[k4B * 4 + i][c4 * 4]);
          acc_gc[i] += v.x * q.x + v.y * q.y + v.z * q.z + v.w * q.w;
        }
      }
    }
  }
  if (gB < G) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k4B * 4 + i < K) atomicAdd(grad_coeff_ws + gB * K + k4B * 4 + i, acc_gc[i]);
  }
}

// ------------------------------------------------------------------------------------------ NMS soft IoU
constexpr int kSiMaxQ = 128;                                   // rows per chunk incl. the all-ones row
constexpr size_t kSiSmem = static_cast<size_t>(2 * kSiMaxQ) * kPad * sizeof(float);

// TRACK = false: NMS form (one logit tensor, rows i = sigmoid, rows j = sigmoid > 0.5, frame stride + nearest 0.5x downsampling).
// TRACK = true:  tracker form (mdqe/tracking/OverTracker.py:92-113): rows i = saved_masks > 0.5, rows j = input_masks > 0.5 on
//                probabilities, every pixel; `mask` / `mask_b` are the two tensors, N2 = T*H*W (T2 = T, H2 = H, W2 = W, t_step = 1).
template <bool TRACK>
__global__ void __launch_bounds__(kThreads)
nms_siou_kernel(const float* __restrict__ mask, const float* __restrict__ mask_b, int i0, int Qi, int j0, int Qj, int T, int H, int W, int t_step,
                int T2, int H2, int W2, int64_t cols_per_cta, float* __restrict__ ws) {
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
      int64_t src = cc;
      if (ok && !TRACK) {                                     // nearest, scale 0.5: source pixel (2y, 2x); frames t_step apart
        const int x2 = static_cast<int>(cc % W2), y2 = static_cast<int>((cc / W2) % H2), t2 = static_cast<int>(cc / (static_cast<int64_t>(W2) * H2));
        src = static_cast<int64_t>(t2) * t_step * plane + static_cast<int64_t>(2 * y2) * W + 2 * x2;
      }
      const int r_end = max(Qi, Qj) + 1;                      // rows beyond stay zero from the initial fill
      for (int r = t >> 5; r < r_end; r += 8) {
        float so = 0.f, ha = 0.f;
        if (ok && TRACK) {
          if (r < Qi) so = __ldg(mask + static_cast<int64_t>(i0 + r) * T * plane + src) > 0.5f ? 1.f : 0.f;       // OverTracker.py:99
          else if (r == Qi) so = 1.f;
          if (r < Qj) ha = __ldg(mask_b + static_cast<int64_t>(j0 + r) * T * plane + src) > 0.5f ? 1.f : 0.f;     // :98
          else if (r == Qj) ha = 1.f;
        } else if (ok) {
          if (r < Qi) so = 1.f / (1.f + expf(-__ldg(mask + static_cast<int64_t>(i0 + r) * T * plane + src)));
          else if (r == Qi) so = 1.f;
          // mask_soft.gt(0.5) on the fp32 sigmoid (mdqe.py:396), not logit > 0: they differ for tiny positive logits
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

__global__ void nms_siou_finalize_kernel(const float* __restrict__ ws, int i0, int Qi, int j0, int Qj, int ld, float eps, float* __restrict__ siou) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Qi * Qj) return;
  const int i = idx / Qj, j = idx % Qj;
  const float num = ws[i * kSiMaxQ + j];
  const float den = ws[i * kSiMaxQ + Qj] + ws[Qi * kSiMaxQ + j] - num;      // mdqe.py:400 / OverTracker.py:106-110
  // eps = 1 (mdqe.py:401) or 1e-6 (OverTracker.py:111; its `saved_valid` factor only zeroes pairs whose numerator is 0 anyway)
  siou[static_cast<int64_t>(i0 + i) * ld + j0 + j] = num / (den + eps);
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


}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

size_t mask_match_cost_workspace_bytes(void) {
  const size_t tc = match_cost_tc_workspace_floats();
  return (tc > static_cast<size_t>(kMcWsFloats) ? tc : static_cast<size_t>(kMcWsFloats)) * sizeof(float);
}

int mask_match_cost(void* stream, const void* coeff, const void* proto, const void* targets, int Q, int K, int G, int64_t Ncols,
                    void* workspace, void* cost_bce, void* cost_dice) {
  if (Q < 0 || G < 0 || K <= 0 || K > 32 || Ncols <= 0) return fail(MSDA_ERR_INVALID_ARG, "mask_match_cost: Q=%d K=%d G=%d Ncols=%lld (K <= 32)", Q, K, G, (long long)Ncols);
  if (Q == 0 || G == 0) return 0;
  if (!coeff || !proto || !targets || !workspace || !cost_bce || !cost_dice) return fail(MSDA_ERR_INVALID_ARG, "mask_match_cost: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tensor-core path (csrc/match_cost_tc.cuh): contraction as 3xTF32 tcgen05 MMAs, costs formed in the epilogue out of TMEM
  const int tc_mode = option("consumer_tc");
  const bool tc_ok = match_cost_tc_eligible(coeff, proto, targets, K, Ncols);
  if (tc_mode == 2 && !tc_ok) return fail(MSDA_ERR_UNSUPPORTED, "mask_match_cost: the tensor-core kernel needs K %% 4 == 0, Ncols %% 4 == 0 and 16-byte aligned tensors");
  if (tc_mode != 1 && tc_ok) {
    for (int q0 = 0; q0 < Q; q0 += 256) {
      const int qn = Q - q0 < 256 ? Q - q0 : 256;
      for (int g0 = 0; g0 < G; g0 += 16) {
        const int gn = G - g0 < 16 ? G - g0 : 16;
        if (int rc = match_cost_tc_dispatch(st, static_cast<const float*>(coeff) + static_cast<int64_t>(q0) * K, static_cast<const float*>(proto),
                                            static_cast<const float*>(targets) + static_cast<int64_t>(g0) * Ncols, qn, K, gn, Ncols,
                                            static_cast<float*>(workspace), static_cast<float*>(cost_bce) + static_cast<int64_t>(q0) * G + g0,
                                            static_cast<float*>(cost_dice) + static_cast<int64_t>(q0) * G + g0, G)) return rc;
      }
    }
    return 0;
  }
  if (int rc = ensure_func_attr(match_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMcSmem))) return rc;
  const int64_t chunks = (Ncols + kTC - 1) / kTC;
  const int slots = option("consumer_ctas") > 0 ? option("consumer_ctas") : 3 * sm_count();      // three resident CTAs per SM
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

static int siou_launch(const char* who, cudaStream_t st, bool track, const float* a, const float* b, int Qa, int Qb, int T, int H, int W,
                       void* workspace, float* siou) {
  if (int rc = ensure_func_attr(nms_siou_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSiSmem))) return rc;
  if (int rc = ensure_func_attr(nms_siou_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSiSmem))) return rc;
  const int t_step = (!track && T >= 5) ? 2 : 1;                // mask_pred[:, ::2] if T >= 5 (mdqe.py:394)
  const int T2 = (T + t_step - 1) / t_step, H2 = track ? H : H / 2, W2 = track ? W : W / 2;
  const int64_t N2 = static_cast<int64_t>(T2) * H2 * W2;
  const int64_t chunks = (N2 + kTC - 1) / kTC;
  // measured (tools/consumer_cta_sweep.py): the sigmoid-heavy NMS form likes two CTAs per SM, the tracker form one
  const int64_t slots = option("consumer_ctas") > 0 ? option("consumer_ctas") : (track ? 1 : 2) * sm_count();
  const int ctas = static_cast<int>(chunks < slots ? chunks : slots);
  const int64_t cols_per_cta = ((chunks + ctas - 1) / ctas) * kTC;
  for (int i0 = 0; i0 < Qa; i0 += kSiMaxQ - 1) {
    const int qi = Qa - i0 < kSiMaxQ - 1 ? Qa - i0 : kSiMaxQ - 1;
    for (int j0 = 0; j0 < Qb; j0 += kSiMaxQ - 1) {
      const int qj = Qb - j0 < kSiMaxQ - 1 ? Qb - j0 : kSiMaxQ - 1;
      if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, mask_nms_siou_workspace_bytes(), st), "cudaMemsetAsync(workspace)")) return rc;
      {
        ProfScope prof(st, 5, static_cast<int64_t>(qi) * N2);
        if (track)
          nms_siou_kernel<true><<<ctas, kThreads, kSiSmem, st>>>(a, b, i0, qi, j0, qj, T, H, W, t_step, T2, H2, W2, cols_per_cta, static_cast<float*>(workspace));
        else
          nms_siou_kernel<false><<<ctas, kThreads, kSiSmem, st>>>(a, b, i0, qi, j0, qj, T, H, W, t_step, T2, H2, W2, cols_per_cta, static_cast<float*>(workspace));
      }
      if (int rc = after_launch(who)) return rc;
      nms_siou_finalize_kernel<<<(qi * qj + 255) / 256, 256, 0, st>>>(static_cast<const float*>(workspace), i0, qi, j0, qj, Qb, track ? 1e-6f : 1.f, siou);
      if (int rc = after_launch("nms_siou_finalize_kernel")) return rc;
    }
  }
  return 0;
}

int mask_nms_siou(void* stream, const void* mask_pred, int Q, int T, int H, int W, void* workspace, void* siou) {
  if (Q < 0 || T <= 0 || H < 2 || W < 2) return fail(MSDA_ERR_INVALID_ARG, "mask_nms_siou: Q=%d T=%d H=%d W=%d", Q, T, H, W);
  if (Q == 0) return 0;
  if (!mask_pred || !workspace || !siou) return fail(MSDA_ERR_INVALID_ARG, "mask_nms_siou: NULL pointer");
  return siou_launch("nms_siou_kernel", static_cast<cudaStream_t>(stream), false, static_cast<const float*>(mask_pred),
                     static_cast<const float*>(mask_pred), Q, Q, T, H, W, workspace, static_cast<float*>(siou));
}

int mask_track_siou(void* stream, const void* saved_masks, const void* input_masks, int Ns, int Ni, int T, int H, int W, void* workspace,
                    void* siou) {
  if (Ns < 0 || Ni < 0 || T <= 0 || H <= 0 || W <= 0) return fail(MSDA_ERR_INVALID_ARG, "mask_track_siou: Ns=%d Ni=%d T=%d H=%d W=%d", Ns, Ni, T, H, W);
  if (Ns == 0 || Ni == 0) return 0;
  if (!saved_masks || !input_masks || !workspace || !siou) return fail(MSDA_ERR_INVALID_ARG, "mask_track_siou: NULL pointer");
  return siou_launch("track_siou_kernel", static_cast<cudaStream_t>(stream), true, static_cast<const float*>(saved_masks),
                     static_cast<const float*>(input_masks), Ns, Ni, T, H, W, workspace, static_cast<float*>(siou));
}

size_t mask_losses_workspace_bytes(void) { return static_cast<size_t>(kMlMaxG) * kMlStats * sizeof(float); }

static int mask_losses_check(const char* who, const void* coeff, const void* proto, const void* tgt, int G, int K, int64_t Ncols) {
  if (G < 0 || G > kMlMaxG || K <= 0 || K > 32 || Ncols <= 0) return fail(MSDA_ERR_INVALID_ARG, "%s: G=%d (<= %d per call) K=%d (<= 32) Ncols=%lld", who, G, kMlMaxG, K, (long long)Ncols);
  if (G > 0 && (!coeff || !proto || !tgt)) return fail(MSDA_ERR_INVALID_ARG, "%s: NULL pointer", who);
  return 0;
}
static void mask_losses_grid(int64_t Ncols, int* ctas, int64_t* cols_per_cta) {
  const int64_t chunks = (Ncols + kTC - 1) / kTC;
  const int slots = 4 * sm_count();
  *ctas = static_cast<int>(chunks < slots ? chunks : slots);
  *cols_per_cta = ((chunks + *ctas - 1) / *ctas) * kTC;
}

int mask_losses_forward(void* stream, const void* coeff, const void* proto, const void* targets, const void* targets_interinst, int G, int K,
                        int64_t Ncols, float num_masks, void* workspace, void* row_stats, void* losses) {
  if (int rc = mask_losses_check("mask_losses_forward", coeff, proto, targets, G, K, Ncols)) return rc;
  if (!workspace || !row_stats || !losses) return fail(MSDA_ERR_INVALID_ARG, "mask_losses_forward: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, mask_losses_workspace_bytes(), st), "cudaMemsetAsync(workspace)")) return rc;
  if (G > 0) {
    int ctas; int64_t cpc;
    mask_losses_grid(Ncols, &ctas, &cpc);
    ProfScope prof(st, 7, static_cast<int64_t>(G) * Ncols);
    mask_losses_fwd_kernel<<<ctas, kThreads, 0, st>>>(static_cast<const float*>(coeff), static_cast<const float*>(proto), static_cast<const float*>(targets),
                                                      static_cast<const float*>(targets_interinst), G, K, Ncols, cpc, static_cast<float*>(workspace));
    if (int rc = after_launch("mask_losses_fwd_kernel")) return rc;
  }
  mask_losses_finalize_kernel<<<1, kMlMaxG, 0, st>>>(static_cast<const float*>(workspace), G, Ncols, targets_interinst != nullptr, num_masks,
                                                     static_cast<float*>(row_stats), static_cast<float*>(losses));
  return after_launch("mask_losses_finalize_kernel");
}

int mask_losses_backward(void* stream, const void* coeff, const void* proto, const void* targets, const void* targets_interinst,
                         const void* row_stats, const void* grad_losses, int G, int K, int64_t Ncols, float num_masks, void* grad_coeff,
                         void* grad_proto) {
  if (int rc = mask_losses_check("mask_losses_backward", coeff, proto, targets, G, K, Ncols)) return rc;
  if (!grad_proto || (G > 0 && (!row_stats || !grad_losses || !grad_coeff))) return fail(MSDA_ERR_INVALID_ARG, "mask_losses_backward: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G == 0) return check_cuda(cudaMemsetAsync(grad_proto, 0, static_cast<size_t>(K) * Ncols * sizeof(float), st), "cudaMemsetAsync(grad_proto)");
  if (int rc = check_cuda(cudaMemsetAsync(grad_coeff, 0, static_cast<size_t>(G) * K * sizeof(float), st), "cudaMemsetAsync(grad_coeff)")) return rc;
  int ctas; int64_t cpc;
  mask_losses_grid(Ncols, &ctas, &cpc);
  {
    ProfScope prof(st, 8, static_cast<int64_t>(G) * Ncols);
    mask_losses_bwd_kernel<<<ctas, kThreads, 0, st>>>(static_cast<const float*>(coeff), static_cast<const float*>(proto), static_cast<const float*>(targets),
                                                      static_cast<const float*>(targets_interinst), static_cast<const float*>(row_stats),
                                                      static_cast<const float*>(grad_losses), G, K, Ncols, cpc, num_masks, static_cast<float*>(grad_coeff),
                                                      static_cast<float*>(grad_proto));
  }
  return after_launch("mask_losses_bwd_kernel");
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
