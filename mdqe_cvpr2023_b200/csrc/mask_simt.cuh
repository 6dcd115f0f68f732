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
template <> __device__ __forceinline__ float mk_ld<__half>(const __half* p) { return __half2float(*p); }

__device__ __forceinline__ void mk_st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void mk_st4(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

__device__ __forceinline__ void mk_st4(__half* p, const float (&v)[4]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
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

// ------------------------------------------------------------------------------------------------------------
// Second-generation fp32 backward: two register-tiled kernels instead of the fused one above (542 us at R50_360;
// its grad_coeff half issued 5 shared loads per 16 FMAs).  grad_out (48 MB) is read by both, back to back, so the
// second read comes out of the 126 MB L2.
//
// grad_proto[b,k,n] = sum_q coeff[b,q,k] * go[b,q,n]:  CTA = 32 k x 256 columns, thread = 4 k x 8 columns,
// q walked in chunks of 28 rows through shared memory (coeff reads are warp broadcasts).
constexpr int kGpTN = 256, kGpQC = 28;
__global__ void __launch_bounds__(256)
mask_grad_proto_kernel(const float* __restrict__ coeff, const float* __restrict__ go, float* __restrict__ gproto,
                       int Q, int K, int64_t Ncols) {
  __shared__ __align__(16) float sC[kGpQC][32];
  __shared__ __align__(16) float sG[kGpQC][kGpTN];
  const int b = blockIdx.z, k0 = blockIdx.y * 32;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kGpTN;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // ty: 4 k each, tx: columns tx*4 and 128 + tx*4
  const float* A = coeff + static_cast<int64_t>(b) * Q * K;
  const float* G = go + static_cast<int64_t>(b) * Q * Ncols;
  const bool vec = (Ncols % 4 == 0);
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int q0 = 0; q0 < Q; q0 += kGpQC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kGpQC * 32; idx += 256) {
      const int qi = idx >> 5, kk = idx & 31;
      sC[qi][kk] = (q0 + qi < Q && k0 + kk < K) ? __ldg(A + static_cast<int64_t>(q0 + qi) * K + k0 + kk) : 0.f;
    }
    for (int idx = threadIdx.x; idx < kGpQC * (kGpTN / 4); idx += 256) {
      const int qi = idx / (kGpTN / 4), c4 = (idx % (kGpTN / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + qi < Q) {
        const float* src = G + static_cast<int64_t>(q0 + qi) * Ncols + n0 + c4;
        if (vec && n0 + c4 + 3 < Ncols) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (n0 + c4 + 0 < Ncols) v.x = __ldg(src + 0);
          if (n0 + c4 + 1 < Ncols) v.y = __ldg(src + 1);
          if (n0 + c4 + 2 < Ncols) v.z = __ldg(src + 2);
          if (n0 + c4 + 3 < Ncols) v.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(&sG[qi][c4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int qi = 0; qi < kGpQC; ++qi) {
      const float4 a = *reinterpret_cast<const float4*>(&sC[qi][ty * 4]);
      const float4 g0 = *reinterpret_cast<const float4*>(&sG[qi][tx * 4]);
      const float4 g1 = *reinterpret_cast<const float4*>(&sG[qi][128 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
    }
  }
  float* GP = gproto + static_cast<int64_t>(b) * K * Ncols;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t n = n0 + h * 128 + tx * 4;
      float* dst = GP + static_cast<int64_t>(k) * Ncols + n;
      if (vec && n + 3 < Ncols) *reinterpret_cast<float4*>(dst) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      else
        for (int j = 0; j < 4; ++j)
          if (n + j < Ncols) dst[j] = acc[i][4 * h + j];
    }
  }
}

// grad_coeff[b,q,k] = sum_n go[b,q,n] * proto[b,k,n]:  CTA = all (<= 256) q x 32 k over a slice of the columns,
// thread = 8 q x 4 k (rows qg + 32 i, columns kg + 8 j: neighbouring lanes touch neighbouring shared-memory rows,
// whose 36-float pitch puts them in different banks), columns walked 32 at a time through shared memory; per-CTA partial sums are added to
// grad_coeff with atomics (the caller zero-fills it).  grid (column slices, ceil(K/32), B * ceil(Q/256)).
constexpr int kGcQ = 256, kGcNC = 32, kGcPitch = kGcNC + 4;
__global__ void __launch_bounds__(256)
mask_grad_coeff_kernel(const float* __restrict__ proto, const float* __restrict__ go, float* __restrict__ gcoeff,
                       int Q, int K, int64_t Ncols, int64_t cols_per_cta, int q_blocks) {
  __shared__ __align__(16) float sG[kGcQ][kGcPitch];      // pitch 36 floats: the 8 q-groups of a warp hit different banks
  __shared__ __align__(16) float sP[32][kGcPitch];
  const int b = blockIdx.z / q_blocks, qb = blockIdx.z % q_blocks;
  const int q_base = qb * kGcQ, k0 = blockIdx.y * 32;
  const int64_t n_begin = static_cast<int64_t>(blockIdx.x) * cols_per_cta;
  const int64_t n_end = min(Ncols, n_begin + cols_per_cta);
  const int kg = threadIdx.x & 7, qg = threadIdx.x >> 3;           // 8 k-groups x 32 q-groups (8 q each)
  const float* Pm = proto + static_cast<int64_t>(b) * K * Ncols;
  const float* G = go + static_cast<int64_t>(b) * Q * Ncols;
  const bool vec = (Ncols % 4 == 0);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t n0 = n_begin; n0 < n_end; n0 += kGcNC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < (kGcQ + 32) * (kGcNC / 4); idx += 256) {
      const int r = idx / (kGcNC / 4), c4 = (idx % (kGcNC / 4)) * 4;
      const bool is_p = r >= kGcQ;
      const int row = is_p ? (r - kGcQ) : r;
      const bool row_ok = is_p ? (k0 + row < K) : (q_base + row < Q);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) {
        const float* src = (is_p ? Pm + static_cast<int64_t>(k0 + row) * Ncols : G + static_cast<int64_t>(q_base + row) * Ncols) + n0 + c4;
        if (vec && n0 + c4 + 3 < n_end) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (n0 + c4 + 0 < n_end) v.x = __ldg(src + 0);
          if (n0 + c4 + 1 < n_end) v.y = __ldg(src + 1);
          if (n0 + c4 + 2 < n_end) v.z = __ldg(src + 2);
          if (n0 + c4 + 3 < n_end) v.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(is_p ? &sP[row][c4] : &sG[row][c4]) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int nn = 0; nn < kGcNC; nn += 4) {
      float4 p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = *reinterpret_cast<const float4*>(&sP[kg + 8 * j][nn]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 g = *reinterpret_cast<const float4*>(&sG[qg + 32 * i][nn]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(g.x, p[j].x, fmaf(g.y, p[j].y, fmaf(g.z, p[j].z, fmaf(g.w, p[j].w, acc[i][j]))));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q_base + qg + 32 * i;
    if (q >= Q) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + kg + 8 * j;
      if (k < K) atomicAdd(gcoeff + (static_cast<int64_t>(b) * Q + q) * K + k, acc[i][j]);
    }
  }
}

}  // namespace msda
