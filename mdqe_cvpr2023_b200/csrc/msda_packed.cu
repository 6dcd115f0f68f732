// bf16 forward on a PAIRED-CORNER value layout (north_star: "bf16x2 loads of the value tensor laid out head-major per level").
//
// What bounds the forward is the number of 128-byte wavefronts on the SM's L1 data pipe: four corner rows per
// sample, one wavefront each -- and a 64-byte bf16 row costs the same wavefront as a 128-byte fp32 row, so bf16 storage in the
// reference layout buys nothing.  Here the value tensor is first re-laid so that the two x-neighbours of a bilinear sample share
// ONE aligned 128-byte line:
//
//   packed[n][prow][m] = { value[n, cell(y, xp - 1), m, 0:32] , value[n, cell(y, xp), m, 0:32] }        (bf16, 64 + 64 bytes)
//   prow = pstart_l + y * (W_l + 1) + xp,   xp in [0, W_l]   (cells outside the level are stored as zeros),
//   pstart_l = sum_{k<l} H_k (W_k + 1);  batch stride = 2 S M lines (an upper bound of sum_l H_l (W_l + 1), known without reading
//   the level table on the host).
//
// A sample with integer corner (x0, y0) then needs the lines (y0, x0 + 1) and (y0 + 1, x0 + 1): TWO wavefronts instead of four,
// and out-of-range x-corners need no predicate (they read zeros).  msda_pack_value builds the layout in one streaming pass
// (value may be fp32 or bf16: it is what value_proj produces); msda_fwd_packed_kernel is the sampler:
//   phase 1 (lane = sample)   geometry, four records {line, weight} per sample: (row 0 | 1) x (left | right half of the line)
//   phase 2 (8 lanes = 1 line) a warp instruction gathers 4 lines = 2 samples; lane (g = lane / 4, c = lane % 4) holds channels
//                             8c..8c+7 of the half-line of record g, its weight comes from one LDS.64 per instruction;
//                             partial sums in 8 registers, folded over the 8 lane groups by a transposing butterfly (7 shuffles
//                             per pair), one store per pair.
// Per pair: 32 line wavefronts + 8 record loads + 4 record stores + 7 shuffles instead of 64 + 17 + 4 + 3.
#include <type_traits>

#include "msda_launch.cuh"

namespace msda {
namespace {

__device__ __forceinline__ FastDiv make_fastdiv_dev(uint32_t d) {      // device twin of make_fastdiv (msda_launch.cuh)
  FastDiv f{d, 0u, 0u};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned p = 31 + l;
  f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  f.shr = p - 32;
  return f;
}

template <typename VT>
__global__ void __launch_bounds__(256)
msda_pack_value_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ level_start,
                       uint4* __restrict__ packed, int N, int S, int M, int L) {
  __shared__ PackedLevel s_lvl[kMaxLevels];
  __shared__ int s_total;
  pdl_wait();
  pdl_trigger();
  stage_packed_levels(s_lvl, &s_total, shapes, level_start, L, S);
  __syncthreads();
  const int Sp = s_total;
  const uint32_t pieces_per_n = static_cast<uint32_t>(Sp) * static_cast<uint32_t>(M) * 8u;      // < 2^32: checked on the host (2 S M lines)
  const int n = blockIdx.y;
  const FastDiv dM = make_fastdiv_dev(static_cast<uint32_t>(M));
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < pieces_per_n; it += gridDim.x * blockDim.x) {
    const int piece = static_cast<int>(it & 7u);
    const uint32_t r = it >> 3;                                                 // line inside the batch element: prow * M + m
    const int prow = static_cast<int>(fd_div(r, dM)), m = static_cast<int>(r - static_cast<uint32_t>(prow) * static_cast<uint32_t>(M));
    int l = 0;
    while (l + 1 < L && prow >= s_lvl[l + 1].pstart) ++l;
    const PackedLevel lv = s_lvl[l];
    const int q = prow - lv.pstart;
    const int y = q / (lv.W + 1), xp = q - y * (lv.W + 1);
    const int x = (piece < 4) ? xp - 1 : xp;                                     // left half: cell xp - 1, right half: cell xp
    const int cb = piece & 3;                                                   // channels 8 cb .. 8 cb + 7
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (x >= 0 && x < lv.W && y < lv.H) {
      const VT* src = value + ((static_cast<int64_t>(n) * S + lv.start + static_cast<int64_t>(y) * lv.W + x) * M + m) * 32 + cb * 8;
      if constexpr (std::is_same<VT, float>::value) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
        o = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
      } else {
        o = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
    packed[(static_cast<int64_t>(n) * 2 * S * M + r) * 8 + piece] = o;      // consecutive threads, consecutive 16-byte pieces
  }
}

struct __align__(8) PRec { uint32_t line; float w; };

__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }

// FUSED: loc / aw are the raw sampling offsets and attention logits of the module (the fused prologue of msda_fast2.cuh: softmax over
// the L*P lanes of a pair, loc = ref + offsets / scale or the decoder's box-scaled grid), optionally as column ranges of ONE query
// projection matrix (fz.row_stride).  OT: bf16, or fp32 for a caller whose next layer (output_proj) takes fp32.
template <typename LT, int LP, bool FUSED, typename OT>
__global__ void __launch_bounds__(kThreads, 6)
msda_fwd_packed_kernel(const uint4* __restrict__ packed, const int64_t* __restrict__ shapes, const int64_t* __restrict__ level_start,
                       const LT* __restrict__ loc,
                       const LT* __restrict__ aw, OT* __restrict__ out, int S, int M, int L, int P, uint32_t n_pairs,
                       int chunk_pairs, FastDiv div_m, FastDiv div_mq, FusedArgs fz) {
  static_assert(!FUSED || std::is_same<LT, float>::value, "the fused prologue takes fp32 offsets / logits");
  constexpr int QPW = 32 / LP;                       // pairs per warp round
  __shared__ PackedLevel s_lvl[kMaxLevels];
  __shared__ int s_total;
  __shared__ __align__(16) PRec s_rec[kWarpsPerCta][QPW * LP * 4];

  pdl_wait();
  pdl_trigger();
  stage_packed_levels(s_lvl, &s_total, shapes, level_start, L, S);      // the same table (and the same disabled levels) as the pack pass
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PRec* my_rec = s_rec[warp];
  const uint32_t chunk_begin = blockIdx.x * static_cast<uint32_t>(chunk_pairs);
  const uint32_t chunk_end = min(n_pairs, chunk_begin + static_cast<uint32_t>(chunk_pairs));
  const int ps = lane / LP, ss = lane - ps * LP;     // phase-1 role (QPW * LP == 32)
  const int lvl = ss / P;
  const uint4* lane_base = packed + (lane & 7);
  const int grp = lane >> 2;                         // record of this lane inside an instruction's 8

  for (uint32_t p0 = chunk_begin + warp * QPW; p0 < chunk_end; p0 += kWarpsPerCta * QPW) {
    const int npair = static_cast<int>(min(static_cast<uint32_t>(QPW), chunk_end - p0));
    // ---- phase 1: one lane per sample
    {
      PRec r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { r[k].line = 0u; r[k].w = 0.f; }
      const bool has_sample = ps < npair;
      float x = 0.f, y = 0.f, a = 0.f;
      uint32_t nq = 0, m = 0, n = 0;
      if (has_sample) {
        const uint32_t pair = p0 + ps;
        nq = fd_div(pair, div_m);
        m = pair - nq * div_m.d;
        n = fd_div(pair, div_mq);
        int64_t li = static_cast<int64_t>(pair) * LP + ss, ai = li;
        if constexpr (FUSED) {
          if (fz.row_stride > 0) {
            const int64_t row = static_cast<int64_t>(nq) * fz.row_stride;
            li = (row >> 1) + m * LP + ss;
            ai = row + m * LP + ss;
          }
        }
        load_loc_aw2<LT>(loc, aw, li, ai, x, y, a);
      }
      if constexpr (FUSED) {                          // every lane of the warp takes part in the softmax shuffles
        a = segment_softmax<LP>(a, has_sample);
        float mk_x, mk_y;
        if (has_sample) fused_location(fz, nq, m, ss, LP, x, y, mk_x, mk_y);
      }
      if (has_sample) {
        const PackedLevel lv = s_lvl[lvl];
        const SampleGeom g = sample_geom(x, y, lv.H, lv.W);
        if (g.x0 >= -1) {                              // sane sample (sample_geom marks the others with -8): x0 + 1 in [0, W]
          const uint32_t stride = static_cast<uint32_t>(lv.W + 1) * static_cast<uint32_t>(M);
          const uint32_t line0 = (n * 2u * static_cast<uint32_t>(S) + static_cast<uint32_t>(lv.pstart)) * static_cast<uint32_t>(M) +
                                 static_cast<uint32_t>(g.y0 * (lv.W + 1) + g.x0 + 1) * static_cast<uint32_t>(M) + m;   // may wrap for y0 = -1: unused then
          const float hx = 1.f - g.lx, hy = 1.f - g.ly;
          if (g.oky0) { r[0].line = line0; r[1].line = line0; r[0].w = hy * hx * a; r[1].w = hy * g.lx * a; }
          if (g.oky1) { r[2].line = line0 + stride; r[3].line = line0 + stride; r[2].w = g.ly * hx * a; r[3].w = g.ly * g.lx * a; }
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(my_rec + lane * 4);            // 32 bytes per lane, consecutive: conflict-free
      dst[0] = make_uint4(r[0].line, __float_as_uint(r[0].w), r[1].line, __float_as_uint(r[1].w));
      dst[1] = make_uint4(r[2].line, __float_as_uint(r[2].w), r[3].line, __float_as_uint(r[3].w));
    }
    __syncwarp();
    // ---- phase 2: 8 lanes per line, 2 samples per instruction
#pragma unroll
    for (int pl = 0; pl < QPW; ++pl) {
      if (pl < npair) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const PRec* stream = my_rec + pl * LP * 4 + grp;
#pragma unroll
        for (int b0 = 0; b0 < LP / 2; b0 += 4) {                           // batches of 4 instructions = 8 samples
          PRec rec[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) rec[u] = stream[(b0 + u) * 8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (rec[u].w != 0.f) {
              const uint4 t = __ldg(lane_base + static_cast<uint64_t>(rec[u].line) * 8u);
              const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
              const float2 ww = make_float2(rec[u].w, rec[u].w);
#pragma unroll
              for (int i = 0; i < 4; ++i) {                                 // bf16 -> fp32 is a shift / a mask; one packed FMA per channel pair
                const float2 r2 = __ffma2_rn(ww, make_float2(__uint_as_float(w4[i] << 16), __uint_as_float(w4[i] & 0xffff0000u)),
                                             make_float2(acc[2 * i], acc[2 * i + 1]));
                acc[2 * i] = r2.x; acc[2 * i + 1] = r2.y;
              }
            }
          }
        }
        // fold over the 8 lane groups (lane bits 2, 3, 4): each stage keeps half of the channels and adds the partner's
        float t4[4], t2[2];
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float keep = b4 ? acc[4 + j] : acc[j], send = b4 ? acc[j] : acc[4 + j];
          t4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float keep = b3 ? t4[2 + j] : t4[j], send = b3 ? t4[j] : t4[2 + j];
          t2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        const float keep = b2 ? t2[1] : t2[0], send = b2 ? t2[0] : t2[1];
        const float v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        const int ch = (lane & 3) * 8 + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
        store_out(out + static_cast<int64_t>(p0 + pl) * 32 + ch, v);
      }
    }
    __syncwarp();
  }
}

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

size_t msda_packed_value_bytes(int N, int S, int M, int D) {
  if (N <= 0 || S <= 0 || M <= 0 || D != 32) return 0;
  return static_cast<size_t>(N) * 2u * static_cast<size_t>(S) * static_cast<size_t>(M) * 128u;
}

int msda_pack_value(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                    int N, int S, int M, int D, int L, void* packed) {
  const char* who = "msda_pack_value";
  if (!value || !shapes || !level_start || !packed) return fail(MSDA_ERR_INVALID_ARG, "%s: NULL pointer", who);
  if (N <= 0 || S <= 0 || M <= 0 || L <= 0) return fail(MSDA_ERR_INVALID_ARG, "%s: non-positive size", who);
  if (D != 32 || L > kMaxLevels) return fail(MSDA_ERR_UNSUPPORTED, "%s: D = %d, L = %d (the paired-corner layout is built for D = 32, L <= %d)", who, D, L, kMaxLevels);
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_BF16_LOC32) return fail(MSDA_ERR_INVALID_ARG, "%s: value must be fp32 or bf16", who);
  if (static_cast<int64_t>(N) * 2 * S * M >= (int64_t(1) << 32)) return fail(MSDA_ERR_UNSUPPORTED, "%s: more than 2^32 lines", who);
  if ((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(packed)) & 15u) return fail(MSDA_ERR_UNSUPPORTED, "%s: tensors must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (static_cast<int64_t>(2) * S * M * 8 >= (int64_t(1) << 31)) return fail(MSDA_ERR_UNSUPPORTED, "%s: batch element too large", who);
  int64_t bx = (static_cast<int64_t>(2) * S * M * 8 + 255) / 256;                // upper bound of the pieces of one batch element
  const int64_t cap = (static_cast<int64_t>(sm_count()) * 16 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 blocks(static_cast<unsigned>(bx), static_cast<unsigned>(N));
  if (dtype == MSDA_F32)
    msda_pack_value_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(value), shapes, level_start,
                                                                                 static_cast<uint4*>(packed), N, S, M, L);
  else
    msda_pack_value_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(value), shapes, level_start,
                                                                                         static_cast<uint4*>(packed), N, S, M, L);
  return after_launch("msda_pack_value_kernel");
}

int msda_forward_packed(void* stream, int dtype, const void* packed, const int64_t* shapes, const int64_t* level_start,
                        const void* loc, const void* aw, int N, int S, int M, int D, int L, int Lq, int P, void* out) {
  const char* who = "msda_forward_packed";
  if (!packed || !shapes || !level_start || !loc || !aw || !out) return fail(MSDA_ERR_INVALID_ARG, "%s: NULL pointer", who);
  if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || Lq <= 0 || P <= 0) return fail(MSDA_ERR_INVALID_ARG, "%s: non-positive size", who);
  if (dtype != MSDA_BF16 && dtype != MSDA_BF16_LOC32) return fail(MSDA_ERR_INVALID_ARG, "%s: dtype must be MSDA_BF16 or MSDA_BF16_LOC32", who);
  if (D != 32 || L * P != 16 || L > kMaxLevels) return fail(MSDA_ERR_UNSUPPORTED, "%s: needs D = 32 and L*P = 16 (got D = %d, L*P = %d)", who, D, L * P);
  const int64_t n_pairs = static_cast<int64_t>(N) * Lq * M;
  if (n_pairs >= (int64_t(1) << 31) / 64 || static_cast<int64_t>(N) * 2 * S * M >= (int64_t(1) << 32)) return fail(MSDA_ERR_UNSUPPORTED, "%s: problem too large for 32-bit indices", who);
  if ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(aw) | reinterpret_cast<uintptr_t>(out)) & 15u)
    return fail(MSDA_ERR_UNSUPPORTED, "%s: tensors must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Problem pb{N, S, M, D, L, Lq, P, n_pairs};
  const int chunk = pick_chunk(pb);
  const unsigned grid = static_cast<unsigned>((n_pairs + chunk - 1) / chunk);
  const FastDiv dm = make_fastdiv(M), dmq = make_fastdiv(static_cast<uint32_t>(M) * static_cast<uint32_t>(Lq));
  ProfScope prof(st, MSDA_PROF_MSDA_FWD, n_pairs);
  const FusedArgs none{nullptr, nullptr, 0, 0, 1.f, 0};
  if (dtype == MSDA_BF16)
    launch_kernel(msda_fwd_packed_kernel<__nv_bfloat16, 16, false, __nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, static_cast<const uint4*>(packed), shapes, level_start,
                  static_cast<const __nv_bfloat16*>(loc), static_cast<const __nv_bfloat16*>(aw), static_cast<__nv_bfloat16*>(out), S, M, L, P,
                  static_cast<uint32_t>(n_pairs), chunk, dm, dmq, none);
  else
    launch_kernel(msda_fwd_packed_kernel<float, 16, false, __nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, static_cast<const uint4*>(packed), shapes, level_start,
                  static_cast<const float*>(loc), static_cast<const float*>(aw), static_cast<__nv_bfloat16*>(out), S, M, L, P,
                  static_cast<uint32_t>(n_pairs), chunk, dm, dmq, none);
  return after_launch("msda_fwd_packed_kernel");
}

int msda_fused_forward_packed_joint(void* stream, const void* packed, const int64_t* shapes, const int64_t* level_start,
                                    const void* ref_points, int R, const void* qproj, int row_stride, const void* grid_, int mode,
                                    float offset_scale, int N, int S, int M, int D, int L, int Lq, int P, int out_dtype, void* out) {
  const char* who = "msda_fused_forward_packed_joint";
  if (!packed || !shapes || !level_start || !ref_points || !qproj || !out) return fail(MSDA_ERR_INVALID_ARG, "%s: NULL pointer", who);
  if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || Lq <= 0 || P <= 0) return fail(MSDA_ERR_INVALID_ARG, "%s: non-positive size", who);
  if (out_dtype != MSDA_F32 && out_dtype != MSDA_BF16) return fail(MSDA_ERR_INVALID_ARG, "%s: out_dtype must be MSDA_F32 or MSDA_BF16", who);
  if (D != 32 || L * P != 16 || L > kMaxLevels) return fail(MSDA_ERR_UNSUPPORTED, "%s: needs D = 32 and L*P = 16 (got D = %d, L*P = %d)", who, D, L * P);
  if ((R != 2 && R != 4) || (mode != 0 && mode != 1) || (mode == 1 && (R != 4 || !grid_)) || !(offset_scale > 0.f))
    return fail(MSDA_ERR_INVALID_ARG, "%s: bad fused arguments (R=%d mode=%d scale=%g)", who, R, mode, (double)offset_scale);
  const int64_t lp = static_cast<int64_t>(M) * L * P;
  if (row_stride < 3 * lp || (row_stride & 3)) return fail(MSDA_ERR_INVALID_ARG, "%s: row_stride=%d must be a multiple of 4 and >= 3*M*L*P", who, row_stride);
  const int64_t n_pairs = static_cast<int64_t>(N) * Lq * M;
  if (n_pairs >= (int64_t(1) << 31) / 64 || static_cast<int64_t>(N) * 2 * S * M >= (int64_t(1) << 32)) return fail(MSDA_ERR_UNSUPPORTED, "%s: problem too large for 32-bit indices", who);
  if ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(qproj) | reinterpret_cast<uintptr_t>(out)) & 15u)
    return fail(MSDA_ERR_UNSUPPORTED, "%s: tensors must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Problem pb{N, S, M, D, L, Lq, P, n_pairs};
  const int chunk = pick_chunk(pb);
  const unsigned grid = static_cast<unsigned>((n_pairs + chunk - 1) / chunk);
  const FastDiv dm = make_fastdiv(M), dmq = make_fastdiv(static_cast<uint32_t>(M) * static_cast<uint32_t>(Lq));
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid_), R, mode, offset_scale, row_stride};
  const float* q = static_cast<const float*>(qproj);
  ProfScope prof(st, MSDA_PROF_MSDA_FWD, n_pairs);
  if (out_dtype == MSDA_F32)
    launch_kernel(msda_fwd_packed_kernel<float, 16, true, float>, dim3(grid), dim3(kThreads), 0, st, static_cast<const uint4*>(packed), shapes, level_start,
                  q, q + 2 * lp, static_cast<float*>(out), S, M, L, P, static_cast<uint32_t>(n_pairs), chunk, dm, dmq, fz);
  else
    launch_kernel(msda_fwd_packed_kernel<float, 16, true, __nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, static_cast<const uint4*>(packed), shapes, level_start,
                  q, q + 2 * lp, static_cast<__nv_bfloat16*>(out), S, M, L, P, static_cast<uint32_t>(n_pairs), chunk, dm, dmq, fz);
  return after_launch("msda_fwd_packed_kernel");
}

}  // extern "C"
