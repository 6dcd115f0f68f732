// Tensor-core gradient of the mask contraction with respect to the per-query coefficients (sm_100a, fp32 via 3xTF32):
//
//   grad_coeff[b, q, k] = sum_n grad_out[b, q, n] * proto[b, k, n]
//   (autograd of einsum 'bqm,bmthw->bqthw', /root/reference/mdqe/models/criterion.py:440, transformer_dec.py:255)
//
// The reduction runs over the plane (n ~ 1e5) and both operands have n contiguous, i.e. both are K-major for the MMA and
// arrive from TMA already in the canonical 128B-swizzle layout: rows of 32 fp32 = one swizzle row.  No transposition, only
// the hi/lo split of 3xTF32 (hi = 19-bit truncation, lo = a - hi; hi*hi + hi*lo + lo*hi accumulates to ~1e-6 of fp32).
//
// One persistent CTA per SM owns a contiguous range of 32-column chunks of one (batch, 256-query block).  Per chunk:
//   warp 0      TMA: grad_out box(es) [128 q][32 n] (x MH halves) + proto box [KP k][32 n]          -> bar_full
//   warps 2-9   split in place (element-wise, so the swizzle is irrelevant): hi stays, lo into the twin tile -> bar_ready
//   warp 1      tcgen05.mma kind::tf32, M = 128 (q), N = KP (k), 4 k-steps x 3 terms x MH halves; commit -> bar_empty
// The [q][k] accumulator lives in TMEM for the whole range (MH x KP columns); at the end warps 2-5 drain it and add it to
// grad_coeff with 16-byte reductions (grad_coeff is zeroed by the dispatcher; every CTA of the block adds its slice).
// The kernel reads grad_out exactly once: algorithmic bytes = (Q + K) * N * 4 per batch item.
#pragma once

#include "mask_tc.cuh"

namespace msda {

constexpr int kGcTcSplitWarps = 8;
constexpr int kGcTcThreads = (2 + kGcTcSplitWarps) * 32;
constexpr int kGcTcMaxStages = 5;
constexpr uint32_t kGcTcHalfBytes = 128u * 128u;                 // 128 rows x 32 fp32

__global__ void __launch_bounds__(kGcTcThreads, 1)
mask_grad_coeff_tc_kernel(const __grid_constant__ CUtensorMap map_go, const __grid_constant__ CUtensorMap map_proto,
                          float* __restrict__ grad_coeff, int Q, int K, int KP, int MH, int n_stages, int n_chunks,
                          int chunks_per_slice, int keep_raw, long long* __restrict__ dbg) {
  const int c_begin = blockIdx.x * chunks_per_slice;
  const int c_end = min(n_chunks, c_begin + chunks_per_slice);
  if (c_begin >= c_end) return;                                  // uniform per CTA, before any barrier / TMEM allocation
  const int b = blockIdx.z;
  const int q_base = blockIdx.y * 128 * MH;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  const uint32_t a_bytes = static_cast<uint32_t>(MH) * kGcTcHalfBytes;
  const uint32_t b_bytes = (static_cast<uint32_t>(KP) * 128u + 1023u) & ~1023u;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;       // [A hi][A lo][B hi][B lo]
  __shared__ __align__(8) uint64_t bars[3 * kGcTcMaxStages + 1];
  __shared__ uint32_t s_tmem_base;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (kGcTcMaxStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * kGcTcMaxStages + s); };
  const uint32_t bar_done = bar0 + 8u * (3 * kGcTcMaxStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), kGcTcSplitWarps); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_go) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_proto) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int c = c_begin; c < c_end; ++c, ++i) {
        const int s = i % n_stages;
        const uint32_t ph = (i / n_stages) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        if (dbg && blockIdx.x == 0 && i < 16) dbg[0 * 16 + i] = clock64();            // TMA issued
        const uint32_t dst = smem_u32(smem) + s * stage_bytes;
        mbar_expect_tx(bar_full(s), a_bytes + static_cast<uint32_t>(KP) * 128u);
        for (int h = 0; h < MH; ++h) tma_load_3d(dst + h * kGcTcHalfBytes, &map_go, bar_full(s), c * 32, q_base + h * 128, b);
        tma_load_3d(dst + 2 * a_bytes, &map_proto, bar_full(s), c * 32, 0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32_m128(static_cast<uint32_t>(KP));
      int i = 0;
      uint32_t acc = 0;
      for (int c = c_begin; c < c_end; ++c, ++i) {
        const int s = i % n_stages;
        const uint32_t ph = (i / n_stages) & 1;
        mbar_wait(bar_ready(s), ph);
        if (dbg && blockIdx.x == 0 && i < 16) dbg[3 * 16 + i] = clock64();            // split done, MMAs issue
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem) + s * stage_bytes, a_lo = a_hi + a_bytes;
        const uint32_t b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
        const uint32_t a_sel[3] = {a_hi, a_hi, a_lo}, b_sel[3] = {b_hi, b_lo, b_hi};       // hi*hi + hi*lo + lo*hi
        for (int h = 0; h < MH; ++h) {
          uint32_t acc_h = acc;
          for (int term = 0; term < 3; ++term)
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t a_desc = umma_desc_sw128(a_sel[term] + h * kGcTcHalfBytes + ks * 32u, 16u, 1024u);
              const uint64_t b_desc = umma_desc_sw128(b_sel[term] + ks * 32u, 16u, 1024u);
              umma_tf32(tmem_base + h * 128u, a_desc, b_desc, idesc, acc_h);
              acc_h = 1;
            }
        }
        acc = 1;
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty(s)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
    }
  } else {
    const int t = threadIdx.x - 64;                               // 0 .. 255
    const uint32_t a_vecs = a_bytes / 16, b_vecs = static_cast<uint32_t>(KP) * 8u;
    int i = 0;
    for (int c = c_begin; c < c_end; ++c, ++i) {
      const int s = i % n_stages;
      const uint32_t ph = (i / n_stages) & 1;
      mbar_wait(bar_full(s), ph);
      if (dbg && blockIdx.x == 0 && t == 0 && i < 16) dbg[1 * 16 + i] = clock64();    // operands landed
      uint8_t* st = smem + s * stage_bytes;
      uint4* a_hi = reinterpret_cast<uint4*>(st);
      uint4* a_lo = reinterpret_cast<uint4*>(st + a_bytes);
      uint4* b_hi = reinterpret_cast<uint4*>(st + 2 * a_bytes);
      uint4* b_lo = reinterpret_cast<uint4*>(st + 2 * a_bytes + b_bytes);
      auto split = [](uint4& v, uint4& lo) {
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
        uint32_t* pl = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t hi = pv[e] & 0xffffe000u;
          pl[e] = __float_as_uint(__uint_as_float(pv[e]) - __uint_as_float(hi));
          pv[e] = hi;
        }
      };
      // The MMA reads only the TF32 bits of an fp32 operand (truncation: checked on the device, tools/mask_keepraw_check.py), so the
      // tile as loaded already is the "hi" operand; only lo = a - trunc(a) has to be produced.
#pragma unroll 4
      for (uint32_t k = t; k < a_vecs; k += kGcTcSplitWarps * 32) { uint4 v = a_hi[k], lo; split(v, lo); if (!keep_raw) a_hi[k] = v; a_lo[k] = lo; }
      for (uint32_t k = t; k < b_vecs; k += kGcTcSplitWarps * 32) { uint4 v = b_hi[k], lo; split(v, lo); if (!keep_raw) b_hi[k] = v; b_lo[k] = lo; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready(s));
      if (dbg && blockIdx.x == 0 && t == 0 && i < 16) dbg[2 * 16 + i] = clock64();    // this warp's split done
    }
    if (warp < 6) {
      // ---- drain: warps 2..5 cover the four TMEM lane quarters (quarter = warp % 4)
      const int quarter = warp & 3;
      mbar_wait(bar_done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int h = 0; h < MH; ++h) {
        const int q = q_base + h * 128 + quarter * 32 + lane;
        float* row = grad_coeff + (static_cast<int64_t>(b) * Q + q) * K;
        const uint32_t lane_base = tmem_base + h * 128u + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 32) {
          float v[32];
          tmem_ld32(lane_base + static_cast<uint32_t>(k0), v);
          if (q < Q) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (k0 + j < K) red_add_f32x4(row + k0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);    // K % 4 == 0
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

}  // namespace msda
