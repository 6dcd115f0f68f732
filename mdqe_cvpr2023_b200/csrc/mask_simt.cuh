// Exact-fp32 SIMT kernels for the mask contraction out[b,q,n] = sum_k coeff[b,q,k] * proto[b,k,n]
// (torch.einsum('bqm,bmthw->bqthw') at /root/reference/mdqe/models/matcher.py:182, criterion.py:440,
// transformer_dec.py:255, mdqe/mdqe.py:384) and its two gradients.
//
// With K = hidden_dim/8 = 32 (24 for Swin-L) the contraction is bound by writing the Q x N output,
// so plain FMA code with register tiling already sits near the HBM roofline in the forward; these
// kernels are also the bit-for-bit-fp32 comparison point for the tensor-core path in mask_gemm.cu.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kMaskTQ = 64;     // query rows per CTA
constexpr int kMaskTN = 128;    // plane columns per CTA
constexpr int kMaskKC = 32;     // k slice held in shared memory

template <typename T> __device__ __forceinline__ float mk_ld(const T* p);
template <> __device__ __forceinline__ float mk_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float mk_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void mk_st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void mk_st4(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// grid: (ceil(Ncols/TN), ceil(Q/TQ), B); block 256 = 8 row groups (8 rows each) x 32 column groups (4 cols each)
template <typename IT, typename OT>
__global__ void __launch_bounds__(256)
mask_fwd_simt_kernel(const IT* __restrict__ coeff, const IT* __restrict__ proto, OT* __restrict__ out,
                     int Q, int K, int64_t Ncols, int vec_ok) {
  __shared__ __align__(16) float sA[kMaskTQ][kMaskKC];
  __shared__ __align__(16) float sB[kMaskKC][kMaskTN];
  const int b = blockIdx.z;
  const int q0 = blockIdx.y * kMaskTQ;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kMaskTN;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const IT* A = coeff + static_cast<int64_t>(b) * Q * K;
  const IT* Bm = proto + static_cast<int64_t>(b) * K * Ncols;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kMaskKC) {
    for (int idx = threadIdx.x; idx < kMaskTQ * kMaskKC; idx += 256) {
      const int qi = idx / kMaskKC, kk = idx % kMaskKC;
      const bool ok = (q0 + qi < Q) && (k0 + kk < K);
      sA[qi][kk] = ok ? mk_ld<IT>(A + static_cast<int64_t>(q0 + qi) * K + k0 + kk) : 0.f;
    }
    for (int idx = threadIdx.x; idx < kMaskKC * kMaskTN; idx += 256) {
      const int kk = idx / kMaskTN, nn = idx % kMaskTN;
      const bool ok = (k0 + kk < K) && (n0 + nn < Ncols);
      sB[kk][nn] = ok ? mk_ld<IT>(Bm + static_cast<int64_t>(k0 + kk) * Ncols + n0 + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kMaskKC; kk += 4) {
      float4 bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) bv[u] = *reinterpret_cast<const float4*>(&sB[kk + u][tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(&sA[ty * 8 + i][kk]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[i][0] = fmaf(av[u], bv[u].x, acc[i][0]);
          acc[i][1] = fmaf(av[u], bv[u].y, acc[i][1]);
          acc[i][2] = fmaf(av[u], bv[u].z, acc[i][2]);
          acc[i][3] = fmaf(av[u], bv[u].w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }

  OT* O = out + static_cast<int64_t>(b) * Q * Ncols;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + ty * 8 + i;
    if (q >= Q) continue;
    if (vec_ok && n0 + tx * 4 + 3 < Ncols) {       // Ncols % 4 == 0 and 16-byte aligned base
      mk_st4(O + static_cast<int64_t>(q) * Ncols + n0 + tx * 4, acc[i]);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < Ncols) st_from_float(O + static_cast<int64_t>(q) * Ncols + n, acc[i][j]);
    }
  }
}

// Backward (fp32 only; the reference trains in fp32, configs/R50_coco.yaml:41-42 AMP disabled).
// grid: (ceil(Ncols/TN), ceil(K/32), B).  One pass over grad_out produces both gradients:
//   grad_proto[b,k,n] = sum_q coeff[b,q,k] go[b,q,n]      (complete per CTA)
//   grad_coeff[b,q,k] = sum_n go[b,q,n] proto[b,k,n]      (per-CTA partial over its n tile -> atomicAdd;
//                                                          the caller zero-fills grad_coeff)
__global__ void __launch_bounds__(256)
mask_bwd_simt_kernel(const float* __restrict__ coeff, const float* __restrict__ proto, const float* __restrict__ go,
                     float* __restrict__ gcoeff, float* __restrict__ gproto, int Q, int K, int64_t Ncols) {
  constexpr int QC = 32;
  __shared__ __align__(16) float sP[kMaskKC][kMaskTN];     // proto tile
  __shared__ __align__(16) float sG[QC][kMaskTN];          // grad_out chunk
  __shared__ __align__(16) float sC[QC][kMaskKC];          // coeff chunk
  const int b = blockIdx.z;
  const int k0 = blockIdx.y * kMaskKC;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kMaskTN;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // grad_proto: 4 k x 4 n per thread
  const int cq = threadIdx.x >> 3, ck = (threadIdx.x & 7) * 4;     // grad_coeff: 1 q x 4 k per thread
  const float* A = coeff + static_cast<int64_t>(b) * Q * K;
  const float* Pm = proto + static_cast<int64_t>(b) * K * Ncols;
  const float* G = go + static_cast<int64_t>(b) * Q * Ncols;

  for (int idx = threadIdx.x; idx < kMaskKC * kMaskTN; idx += 256) {
    const int kk = idx / kMaskTN, nn = idx % kMaskTN;
    const bool ok = (k0 + kk < K) && (n0 + nn < Ncols);
    sP[kk][nn] = ok ? __ldg(Pm + static_cast<int64_t>(k0 + kk) * Ncols + n0 + nn) : 0.f;
  }
  float gp[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) gp[i][j] = 0.f;

  for (int q0 = 0; q0 < Q; q0 += QC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < QC * kMaskTN; idx += 256) {
      const int qi = idx / kMaskTN, nn = idx % kMaskTN;
      const bool ok = (q0 + qi < Q) && (n0 + nn < Ncols);
      sG[qi][nn] = ok ? __ldg(G + static_cast<int64_t>(q0 + qi) * Ncols + n0 + nn) : 0.f;
    }
    for (int idx = threadIdx.x; idx < QC * kMaskKC; idx += 256) {
      const int qi = idx / kMaskKC, kk = idx % kMaskKC;
      const bool ok = (q0 + qi < Q) && (k0 + kk < K);
      sC[qi][kk] = ok ? __ldg(A + static_cast<int64_t>(q0 + qi) * K + k0 + kk) : 0.f;
    }
    __syncthreads();
    if (gproto) {
#pragma unroll 8
      for (int qi = 0; qi < QC; ++qi) {
        const float4 a = *reinterpret_cast<const float4*>(&sC[qi][ty * 4]);
        const float4 g = *reinterpret_cast<const float4*>(&sG[qi][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          gp[i][0] = fmaf(av[i], g.x, gp[i][0]);
          gp[i][1] = fmaf(av[i], g.y, gp[i][1]);
          gp[i][2] = fmaf(av[i], g.z, gp[i][2]);
          gp[i][3] = fmaf(av[i], g.w, gp[i][3]);
        }
      }
    }
    if (gcoeff) {
      float gc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int nn = 0; nn < kMaskTN; nn += 4) {
        const float4 g = *reinterpret_cast<const float4*>(&sG[cq][nn]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 p = *reinterpret_cast<const float4*>(&sP[ck + i][nn]);
          gc[i] = fmaf(g.x, p.x, fmaf(g.y, p.y, fmaf(g.z, p.z, fmaf(g.w, p.w, gc[i]))));
        }
      }
      if (q0 + cq < Q) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + ck + i < K) atomicAdd(gcoeff + (static_cast<int64_t>(b) * Q + q0 + cq) * K + k0 + ck + i, gc[i]);
      }
    }
  }
  if (gproto) {
    float* GP = gproto + static_cast<int64_t>(b) * K * Ncols;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty * 4 + i;
      if (k >= K) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n < Ncols) GP[static_cast<int64_t>(k) * Ncols + n] = gp[i][j];
      }
    }
  }
}

}  // namespace msda
