// Kernel selection and launch of the sampling kernels, one instantiation per dtype combination.  The templates below are
// instantiated explicitly in msda_launch_{f32,bf16,bf16_loc32,f64}.cu (four translation units that nvcc compiles in parallel,
// mdqe_cvpr2023_b200/build.py) and declared `extern` for msda_api.cu, which holds the C ABI.
#pragma once

#include <cstdlib>
#include <type_traits>

#include "msda_fast.cuh"
#include "msda_fast2.cuh"
#include "msda_generic.cuh"
#include "msda_internal.h"

#ifndef MSDA_HEAD_RUN_DEFAULT
#define MSDA_HEAD_RUN_DEFAULT 1
#endif

namespace msda {

struct Problem {
  int N, S, M, D, L, Lq, P;
  int64_t n_pairs;
  int G = 1;            // level-table groups sharing loc/aw (temporal form); 1 = the plain operator
  float scale = 1.f;    // out = scale * sum over groups
  FusedArgs fz{nullptr, nullptr, 0, 0, 1.f};   // fused softmax / location prologue (ref != nullptr)
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline bool fast_eligible(int dtype, const Problem& pb, const void* a, const void* b, const void* c) {
  if (dtype == MSDA_F64) return false;
  if (pb.D != 32 && pb.D != 24) return false;
  if (pb.L > kMaxLevels || pb.L * pb.P > 32) return false;
  if ((int64_t)pb.N * pb.S * pb.M * pb.D >= (int64_t(1) << 31)) return false;   // 32-bit row offsets
  if (pb.n_pairs >= (int64_t(1) << 31) / 64) return false;                        // 32-bit pair / sample indices
  return aligned16(a) && aligned16(b) && aligned16(c);
}

inline int pick_chunk(const Problem& pb) {
  const int qpw = 32 / (pb.L * pb.P);
  const int unit = kWarpsPerCta * qpw;                   // pairs one CTA round covers
  int chunk = options().chunk_pairs.load();
  if (chunk <= 0) {
    const int64_t want = pb.n_pairs / (int64_t(sm_count()) * 8);  // aim at >= 8 CTAs per SM
    chunk = static_cast<int>(want < 48 ? want : 48);        // 48: 3400 CTAs on the encoder shape = 5.7 waves of 4 CTAs/SM (64: 4.3 waves, 2 % slower)
  }
  chunk = ((chunk + unit - 1) / unit) * unit;
  return chunk < unit ? unit : chunk;
}

// Head-run pair order (PairMap in msda_fast2.cuh): on for calls large enough that co-resident CTAs would otherwise sit in
// unrelated parts of the image.  The grid becomes ceil(N*Lq / chunk) runs x M heads.
// Measured (profiles/r02_pair_map.md): the backward gains 3-6 % on every encoder shape (L1 hit rate 28 -> 41 %, L2 read sectors
// 20.1 M -> 12.1 M per call); the forward's hit rate rises too (43 -> 61 %) but its time does not move (it is bound by the number
// of wavefronts, not by where the rows come from), so `auto` keeps the forward on the linear order.
inline bool head_run(const Problem& pb, bool split, bool backward) {
  const int mode = options().pair_map.load();
  if (split || mode == 1) return false;
  if (mode == 2) return true;
  return MSDA_HEAD_RUN_DEFAULT != 0 && backward && pb.n_pairs >= 148LL * 4 * 64;
}
inline unsigned head_run_grid(const Problem& pb, int chunk) {
  const int64_t nq = pb.n_pairs / pb.M;
  return static_cast<unsigned>(((nq + chunk - 1) / chunk) * pb.M);
}

inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                                   // ceil(log2 d)
  const unsigned p = 31 + l;
  f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  f.shr = p - 32;
  return f;
}

inline bool fast2_lp(int lp) { return lp == 16 || lp == 12 || lp == 8; }

// Split the level tables of a grouped call across CTAs?  Only worth it (and only implemented) for all-fp32 calls
// that cannot fill the GPU on their own: fewer than ~4 CTAs per SM.
template <typename VT, typename LT>
inline bool split_groups(const Problem& pb) {
  if (pb.G <= 1 || !std::is_same<VT, float>::value || !std::is_same<LT, float>::value) return false;
  if (pb.fz.ref != nullptr) return false;                   // the fused softmax backward needs the sum over all tables
  if (options().chunk_pairs.load() < 0) return false;           // chunk_pairs = -1 disables the split (A/B timing)
  return pb.n_pairs < 148LL * 4 * 16;
}

// The driver's default shared-memory carve-out for the sampling kernels is 132 KB (ncu launch__shared_mem_config_size,
// profiles/r01z) although their resident CTAs need < 64 KB, which leaves L1 only ~120 KB of the SM's 256 KB.  Asking for exactly
// what `ctas_per_sm` resident CTAs use (static + 1 KB reserved each) is worth 1-2 % on every backward shape (183 -> 181 us at 360p,
// 564 -> 560 us at 720p) and nothing on the forward (82 -> 84 us: left at the driver's choice) -- the reuse distance of the
// gathered rows is beyond either L1 size.  Once per instantiation; MSDA_DEFAULT_CARVEOUT=1 keeps the driver's choice (A/B).
template <typename K>
inline bool prefer_small_carveout(K kernel, int ctas_per_sm) {
  const char* env = getenv("MSDA_DEFAULT_CARVEOUT");
  if (env && env[0] == '1') return false;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); return false; }
  const size_t need = static_cast<size_t>(ctas_per_sm) * (fa.sharedSizeBytes + 1024);
  int pct = static_cast<int>((need * 100 + 233471) / 233472);
  if (pct > 100) pct = 100;
  return ensure_func_attr(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) == 0;      // remembered per device
}

// Programmatic dependent launch (option "pdl", default on): the sampling kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs may become resident while this kernel's last
// wave drains and the launch / CTA-scheduling latency between two back-to-back kernels of a step (74 launches per clip, 48 of
// them only 5-25 us long) is hidden.  Every kernel executes `griddepcontrol.wait` before it touches memory (pdl_wait() is
// its first statement), which blocks until the preceding kernel has completed and flushed, so ordering is unchanged.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = options().pdl.load() != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifndef MSDA_FWD_MINB
#define MSDA_FWD_MINB 6                      // resident CTAs per SM promised to ptxas for the register-lean forward (A/B: tools/fwd_variants.sh)
#endif

template <typename VT, typename LT, int D, int MINB>
inline void launch_fwd2_lp(cudaStream_t st, const Problem& pb, dim3 grid, int chunk, int hrun, const VT* v, const int64_t* shapes,
                           const int64_t* lsi, const LT* lc, const LT* a, VT* o) {
  const FastDiv dm = make_fastdiv(pb.M), dmq = make_fastdiv((uint32_t)pb.M * (uint32_t)pb.Lq);
  const uint32_t np = static_cast<uint32_t>(pb.n_pairs);
#define MSDA_FWD2_(LPV, GRP, FUS) launch_kernel(msda_fwd_fast2_kernel<VT, LT, D, LPV, MINB, GRP, FUS>, grid, dim3(kThreads), 0, st, \
      v, shapes, lsi, lc, a, o, pb.S, pb.M, pb.L, pb.P, np, chunk, dm, dmq, pb.G, pb.scale, pb.fz, hrun)
  // the fused prologue (fp32, L*P a power of two: fused_supported) is its own instantiation of the kernels
#define MSDA_FWD2(LPV, GRP) do { \
      if constexpr (std::is_same<VT, float>::value && std::is_same<LT, float>::value && ((LPV) & ((LPV) - 1)) == 0) { \
        if (pb.fz.ref != nullptr) { MSDA_FWD2_(LPV, GRP, true); break; } \
      } \
      MSDA_FWD2_(LPV, GRP, false); } while (0)
  const bool grouped = pb.G > 1 || pb.scale != 1.f;
  switch (pb.L * pb.P) {
    case 16: if (grouped) MSDA_FWD2(16, true); else MSDA_FWD2(16, false); break;
    case 12: if (grouped) MSDA_FWD2(12, true); else MSDA_FWD2(12, false); break;
    default: if (grouped) MSDA_FWD2(8, true); else MSDA_FWD2(8, false); break;
  }
#undef MSDA_FWD2_
#undef MSDA_FWD2
}

template <typename VT, typename LT, int D>
inline void launch_bwd2_lp(cudaStream_t st, const Problem& pb, dim3 grid, int chunk, int hrun, const VT* v, const int64_t* shapes,
                           const int64_t* lsi, const LT* lc, const LT* a, const VT* go, float* gv, LT* gl, LT* ga) {
  const FastDiv dm = make_fastdiv(pb.M), dmq = make_fastdiv((uint32_t)pb.M * (uint32_t)pb.Lq);
  const uint32_t np = static_cast<uint32_t>(pb.n_pairs);
#define MSDA_BWD2_(LPV, GRP, FUS) do { \
      prefer_small_carveout(msda_bwd_fast2_kernel<VT, LT, D, LPV, GRP, FUS>, (GRP) ? 2 : MSDA_BWD_MINB); \
      launch_kernel(msda_bwd_fast2_kernel<VT, LT, D, LPV, GRP, FUS>, grid, dim3(kThreads), 0, st, \
      v, shapes, lsi, lc, a, go, gv, gl, ga, pb.S, pb.M, pb.L, pb.P, np, chunk, dm, dmq, pb.G, pb.scale, pb.fz, merge, hrun); } while (0)
#define MSDA_BWD2(LPV, GRP) do { \
      if constexpr (std::is_same<VT, float>::value && std::is_same<LT, float>::value && ((LPV) & ((LPV) - 1)) == 0) { \
        if (pb.fz.ref != nullptr) { MSDA_BWD2_(LPV, GRP, true); break; } \
      } \
      MSDA_BWD2_(LPV, GRP, false); } while (0)
  const bool grouped = pb.G > 1 || pb.scale != 1.f;
  const int merge = (options().bwd_merge.load() != 0 && (pb.P == 4 || pb.P == 2)) ? pb.P : 0;
  switch (pb.L * pb.P) {
    case 16: if (grouped) MSDA_BWD2(16, true); else MSDA_BWD2(16, false); break;
    case 12: if (grouped) MSDA_BWD2(12, true); else MSDA_BWD2(12, false); break;
    default: if (grouped) MSDA_BWD2(8, true); else MSDA_BWD2(8, false); break;
  }
#undef MSDA_BWD2
#undef MSDA_BWD2_
}

template <typename VT, typename LT>
int launch_fwd(cudaStream_t st, const Problem& pb, bool fast, const void* value, const int64_t* shapes,
                      const int64_t* lsi, const void* loc, const void* aw, void* out) {
  const VT* v = static_cast<const VT*>(value);
  const LT* lc = static_cast<const LT*>(loc);
  const LT* a = static_cast<const LT*>(aw);
  VT* o = static_cast<VT*>(out);
  ProfScope prof(st, MSDA_PROF_MSDA_FWD, pb.n_pairs);
  if constexpr (!std::is_same<VT, double>::value) {
    if (fast) {
      const int chunk = pick_chunk(pb);
      const unsigned grid = static_cast<unsigned>((pb.n_pairs + chunk - 1) / chunk);
      if (options().fwd_variant.load() != 2 && fast2_lp(pb.L * pb.P)) {
        dim3 grid2(grid, 1, 1);
        if (split_groups<VT, LT>(pb)) {      // small grouped call: one CTA row per level table, reductions into zeroed `out`
          grid2.y = pb.G;
          if (check_cuda(cudaMemsetAsync(o, 0, (size_t)pb.n_pairs * pb.D * sizeof(VT), st), "cudaMemsetAsync(out)")) return MSDA_ERR_CUDA;
        }
        // default: register-lean schedule (82 vs 90 us on the encoder shape, profiles/r01d); 3 = batched gathers.  Calls too small
        // to fill the SMs (the decoder's 196 queries: 392 CTAs) are latency bound and take the batched build (8.2 -> 7.4 us, cold)
        const int hrun = head_run(pb, grid2.y > 1, false) ? 1 : 0;
        if (hrun) grid2.x = head_run_grid(pb, chunk);
        const bool small_grid = options().fwd_variant.load() == 0 && static_cast<uint64_t>(grid) * grid2.y < 148u * 4u;
        const bool lean = options().fwd_variant.load() != 3 && !small_grid;
        if (pb.D == 32) {
          if (lean) launch_fwd2_lp<VT, LT, 32, MSDA_FWD_MINB>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, o);
          else launch_fwd2_lp<VT, LT, 32, 3>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, o);
        } else {
          if (lean) launch_fwd2_lp<VT, LT, 24, MSDA_FWD_MINB>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, o);
          else launch_fwd2_lp<VT, LT, 24, 3>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, o);
        }
        return after_launch("msda_fwd_fast2_kernel");
      }
      if (pb.D == 32)
        msda_fwd_fast_kernel<VT, LT, 32><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, o, pb.S, pb.M, pb.L, pb.Lq, pb.P, pb.n_pairs, chunk);
      else
        msda_fwd_fast_kernel<VT, LT, 24><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, o, pb.S, pb.M, pb.L, pb.Lq, pb.P, pb.n_pairs, chunk);
      return after_launch("msda_fwd_fast_kernel");
    }
  }
  const int64_t blocks = (pb.n_pairs + kWarpsPerCta - 1) / kWarpsPerCta;
  const unsigned grid = static_cast<unsigned>(blocks < 148 * 64 ? blocks : 148 * 64);
  msda_fwd_generic_kernel<VT, LT><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, o, pb.S, pb.M, pb.D, pb.L, pb.Lq, pb.P, pb.n_pairs);
  return after_launch("msda_fwd_generic_kernel");
}

template <typename VT, typename LT, typename GT>
int launch_bwd(cudaStream_t st, const Problem& pb, bool fast, const void* value, const int64_t* shapes,
                      const int64_t* lsi, const void* loc, const void* aw, const void* grad_out, GT* gv_acc,
                      void* grad_loc, void* grad_aw) {
  const VT* v = static_cast<const VT*>(value);
  const LT* lc = static_cast<const LT*>(loc);
  const LT* a = static_cast<const LT*>(aw);
  const VT* go = static_cast<const VT*>(grad_out);
  LT* gl = static_cast<LT*>(grad_loc);
  LT* ga = static_cast<LT*>(grad_aw);
  ProfScope prof(st, MSDA_PROF_MSDA_BWD, pb.n_pairs);
  if constexpr (std::is_same<GT, float>::value) {
    if (fast) {
      const int chunk = pick_chunk(pb);
      const unsigned grid = static_cast<unsigned>((pb.n_pairs + chunk - 1) / chunk);
      if (options().bwd_variant.load() != 2 && fast2_lp(pb.L * pb.P)) {
        dim3 grid2(grid, 1, 1);
        if (split_groups<VT, LT>(pb)) {
          grid2.y = pb.G;
          const size_t n_smp = (size_t)pb.n_pairs * pb.L * pb.P;
          if (check_cuda(cudaMemsetAsync(gl, 0, n_smp * 2 * sizeof(LT), st), "cudaMemsetAsync(grad_loc)")) return MSDA_ERR_CUDA;
          if (check_cuda(cudaMemsetAsync(ga, 0, n_smp * sizeof(LT), st), "cudaMemsetAsync(grad_aw)")) return MSDA_ERR_CUDA;
        }
        const int hrun = head_run(pb, grid2.y > 1, true) ? 1 : 0;
        if (hrun) grid2.x = head_run_grid(pb, chunk);
        if (pb.D == 32) launch_bwd2_lp<VT, LT, 32>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, go, gv_acc, gl, ga);
        else launch_bwd2_lp<VT, LT, 24>(st, pb, grid2, chunk, hrun, v, shapes, lsi, lc, a, go, gv_acc, gl, ga);
        return after_launch("msda_bwd_fast2_kernel");
      }
      if (pb.D == 32)
        msda_bwd_fast_kernel<VT, LT, 32><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, go, gv_acc, gl, ga, pb.S, pb.M, pb.L, pb.Lq, pb.P, pb.n_pairs, chunk);
      else
        msda_bwd_fast_kernel<VT, LT, 24><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, go, gv_acc, gl, ga, pb.S, pb.M, pb.L, pb.Lq, pb.P, pb.n_pairs, chunk);
      return after_launch("msda_bwd_fast_kernel");
    }
  }
  const int64_t blocks = (pb.n_pairs + kWarpsPerCta - 1) / kWarpsPerCta;
  const unsigned grid = static_cast<unsigned>(blocks < 148 * 64 ? blocks : 148 * 64);
  msda_bwd_generic_kernel<VT, LT, GT><<<grid, kThreads, 0, st>>>(v, shapes, lsi, lc, a, go, gv_acc, gl, ga, pb.S, pb.M, pb.D, pb.L, pb.Lq, pb.P, pb.n_pairs);
  return after_launch("msda_bwd_generic_kernel");
}


#define MSDA_LAUNCH_EXTERN(KW, VT, LT, GT)                                                                                       \
  KW template int launch_fwd<VT, LT>(cudaStream_t, const Problem&, bool, const void*, const int64_t*, const int64_t*, const void*, \
                                     const void*, void*);                                                                        \
  KW template int launch_bwd<VT, LT, GT>(cudaStream_t, const Problem&, bool, const void*, const int64_t*, const int64_t*,       \
                                         const void*, const void*, const void*, GT*, void*, void*);

}  // namespace msda
