// Mask contraction dispatch: out[b,q,n] = sum_k coeff[b,q,k] * proto[b,k,n] and its gradients.
// See include/msda_b200.h (mask_logits_*) for the contract.
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "mask_simt.cuh"
#include "mask_tc.cuh"
#include "mask_tc4.cuh"
#include "mask_tc_bwd.cuh"
#include "gemm3x.cuh"
#include "match_cost_tc.cuh"
#include "msda_internal.h"
#include "msda_launch.cuh"

namespace msda {

// ------------------------------------------------------------------------------ tcgen05 / TMA path
// cuTensorMapEncodeTiled is a driver-API entry; it is resolved through the runtime so that the library
// does not link against libcuda (and still loads on a box without a driver, e.g. for the CPU-side tests).
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// bf16 / fp32 tensor [d2, d1, d0] (d0 contiguous) -> 3-D tiled map with a {128 bytes, box1, 1} box and 128B swizzle.
static int make_map_in(CUtensorMap* map, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box1) {
  const bool fp32 = dtype == MSDA_F32;
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const uint64_t es = fp32 ? 4 : 2;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * es, d0 * d1 * es};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / es), box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == MSDA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         option("mask_debug") == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return 0;
}

static bool mask_tc_eligible(int in_dtype, const void* coeff, const void* proto, int Q, int K, int64_t Ncols) {
  if (in_dtype != MSDA_BF16 && in_dtype != MSDA_F16) return false;
  if (K < 8 || K > 64 || K % 8 != 0) return false;                     // 16-byte global strides, <= 4 K steps
  if (Q < 1) return false;
  if (Ncols % 8 != 0 || Ncols >= (int64_t(1) << 31)) return false;
  return ((reinterpret_cast<uintptr_t>(coeff) | reinterpret_cast<uintptr_t>(proto)) & 15u) == 0;
}

static long long* g_mask_dbg = nullptr;
// debugging aid: per-item clock64 stamps of CTA 0 (rows: TMA issued, MMA waits, operands landed, epilogue starts, epilogue ends)
int mask_debug_copy(long long* host80) {
  if (!g_mask_dbg) return fail(MSDA_ERR_INVALID_ARG, "mask_debug: no debug buffer (set option mask_debug=1 and run the tcgen05 kernel)");
  return check_cuda(cudaMemcpy(host80, g_mask_dbg, 80 * sizeof(long long), cudaMemcpyDeviceToHost), "cudaMemcpy(mask_debug)");
}

// Persistent pipelined kernel.  The query range is cut into chunks (multiples of 16 rows, <= 256) so that there
// are at least ~6 work items per SM: with whole-Q items a 360p clip has only 480 tiles for 148 SMs (3.24 per SM,
// i.e. a 4-vs-3 imbalance).
// [d2, d1, d0] tensor of 2- or 4-byte elements, no swizzle, box {128, 32, 1}: the output side of the epilogue.
static int make_map_out(CUtensorMap* map, void* base, bool bf16, uint64_t d0, uint64_t d1, uint64_t d2) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const uint64_t es = bf16 ? 2 : 4;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * es, d0 * d1 * es};
  const cuuint32_t box[3] = {kTcTileN, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed (CUresult %d)", static_cast<int>(r));
  return 0;
}

template <typename OT>
static int launch_mask_tc2(cudaStream_t st, int in_dtype, const void* coeff, const void* proto, void* out, int B, int Q, int K,
                           int64_t Ncols) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int KP = (K + 15) / 16 * 16;
  const int n_tiles_n = static_cast<int>((Ncols + kTcTileN - 1) / kTcTileN);
  const int64_t tiles = (int64_t)B * n_tiles_n;
  // query chunks: starts at multiples of 32 (QS), the last chunk takes the remainder
  int n_qchunks = (Q + 255) / 256;                                      // at most 256 query rows per item (one MMA N extent)
  while (n_qchunks < 4 && tiles * n_qchunks < 6LL * sms && (Q / (n_qchunks + 1)) / 32 * 32 >= 32) ++n_qchunks;
  int QS = 0, QN = 0;
  for (;; ++n_qchunks) {
    QS = (n_qchunks == 1) ? ((Q + 31) / 32 * 32) : (Q / n_qchunks) / 32 * 32;
    if (QS < 32) break;
    const int last_rows = Q - (n_qchunks - 1) * QS;
    QN = ((QS > last_rows ? QS : last_rows) + 15) / 16 * 16;
    if (QN <= 256 && mask_tc2_smem_bytes(KP, QN, sizeof(OT)) <= 220 * 1024) break;
  }
  if (QN > 256 || QS < 32) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_forward: cannot chunk Q=%d for the tensor-core kernel", Q);
  const int64_t n_items = tiles * n_qchunks;
  if (n_items >= (int64_t(1) << 31)) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_forward: too many tiles");
  CUtensorMap map_proto, map_coeff, map_out;
  if (int rc = make_map_in(&map_proto, proto, in_dtype, (uint64_t)Ncols, (uint64_t)K, (uint64_t)B, (uint32_t)KP)) return rc;
  if (int rc = make_map_in(&map_coeff, coeff, in_dtype, (uint64_t)K, (uint64_t)Q, (uint64_t)B, (uint32_t)QN)) return rc;
  if (int rc = make_map_out(&map_out, out, sizeof(OT) == 2, (uint64_t)Ncols, (uint64_t)Q, (uint64_t)B)) return rc;
  const size_t smem = mask_tc2_smem_bytes(KP, QN, sizeof(OT));
  if (int rc = ensure_func_attr(mask_fwd_tc2_kernel<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) return rc;
  const unsigned grid = static_cast<unsigned>(n_items < sms ? n_items : sms);
  long long* dbg = nullptr;
  if (option("mask_debug")) {
    static long long* d_dbg = nullptr;
    if (!d_dbg) cudaMalloc(&d_dbg, 5 * 16 * sizeof(long long));
    cudaMemsetAsync(d_dbg, 0, 5 * 16 * sizeof(long long), st);
    dbg = d_dbg;
    g_mask_dbg = d_dbg;
  }
  ProfScope prof(st, MSDA_PROF_MASK_FWD, (int64_t)B * Q * Ncols);
  mask_fwd_tc2_kernel<OT><<<grid, kTc2Threads, smem, st>>>(map_proto, map_coeff, map_out, Q, KP, QS, QN, n_qchunks, n_tiles_n,
                                                          static_cast<int>(n_items), in_dtype == MSDA_F16 ? 1 : 0, dbg);
  return after_launch("mask_fwd_tc2_kernel");
}

// fp32 inputs: 3xTF32 on the tensor cores (mask_fwd_tc4_kernel)
static bool mask_tc3_eligible(int in_dtype, const void* coeff, const void* proto, int Q, int K, int64_t Ncols) {
  if (in_dtype != MSDA_F32) return false;
  if (K < 4 || K % 4 != 0) return false;                               // 16-byte global row strides (reduction walked 32 at a time)
  if (Q < 1) return false;
  if (Ncols % 4 != 0 || Ncols >= (int64_t(1) << 31)) return false;     // 16-byte global strides
  return ((reinterpret_cast<uintptr_t>(coeff) | reinterpret_cast<uintptr_t>(proto)) & 15u) == 0;
}

// fp32 tensor [d2, d1, d0], box {32, 32, 1}, "128B swizzle with 32B atoms": the MN-major tf32 operand layout (UMMA layout type 1)
static int make_map_mn_f32(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         option("mask_debug") == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled(MN-major fp32) failed (CUresult %d)", static_cast<int>(r));
  return 0;
}

// out[b, r, n] = sum_k A[b, r, k] * P[b, k, n]  (kTransB: A is given as [b, k, r]).  Q = rows r, K = reduction length.
template <typename OT, bool kTransB>
static int launch_mask_tc4(cudaStream_t st, const void* coeff, const void* proto, void* out, int B, int Q, int K,
                           int64_t Ncols, int prof_kind = MSDA_PROF_MASK_FWD) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_kchunks = (K + 31) / 32;
  const int n_tiles_n = static_cast<int>((Ncols + kTcTileN - 1) / kTcTileN);
  const int64_t tiles = (int64_t)B * n_tiles_n;
  int n_qchunks = (Q + 127) / 128;
  while (n_qchunks < 4 && tiles * n_qchunks < 6LL * sms && (Q / (n_qchunks + 1)) / 32 * 32 >= 32) ++n_qchunks;
  int QS = 0, QN = 0;
  for (;; ++n_qchunks) {
    QS = (n_qchunks == 1) ? ((Q + 31) / 32 * 32) : (Q / n_qchunks) / 32 * 32;
    if (QS < 32) break;
    const int last_rows = Q - (n_qchunks - 1) * QS;
    QN = ((QS > last_rows ? QS : last_rows) + 15) / 16 * 16;
    if (QN <= 128) break;
  }
  if (QN > 128 || QS < 32) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits: cannot chunk %d rows for the tensor-core kernel", Q);
  const int64_t n_items = tiles * n_qchunks;
  if (n_items >= (int64_t(1) << 31)) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits: too many tiles");
  CUtensorMap map_plane, map_rows, map_out;
  if (int rc = make_map_mn_f32(&map_plane, proto, (uint64_t)Ncols, (uint64_t)K, (uint64_t)B)) return rc;
  if (kTransB) {
    if (int rc = make_map_mn_f32(&map_rows, coeff, (uint64_t)Q, (uint64_t)K, (uint64_t)B)) return rc;
  } else {
    if (int rc = make_map_in(&map_rows, coeff, MSDA_F32, (uint64_t)K, (uint64_t)Q, (uint64_t)B, (uint32_t)QN)) return rc;
  }
  if (int rc = make_map_out(&map_out, out, sizeof(OT) == 2, (uint64_t)Ncols, (uint64_t)Q, (uint64_t)B)) return rc;
  // plane operand through tensor memory (mask_tc4.cuh; a compile-time variant of the kernel): option mask_a_tmem 1 = off (A/B).
  // tools/mask_atm_ab.py, profiles/r02aw_mask_atm_ab.txt (graph replays, same box): forward 13.6 -> 12.0 us at R50_360, 40.4 -> 35.9 at
  // R50_720, 37.1 -> 32.5 at Q = 300; grad_proto (seven chunks of N = 32 MMAs per tile) 23.2 -> 22.8, 53.1 -> 52.1
  const bool a_tm = option("mask_a_tmem") != 1;
  const size_t stage = mask_tc4_stage_bytes(QN, kTransB, a_tm);
  const size_t out_bytes = 4 * 32 * kTcTileN * sizeof(OT);
  int n_stages = static_cast<int>((224 * 1024 - out_bytes) / stage);
  if (n_stages > kTc4MaxStages) n_stages = kTc4MaxStages;
  if (n_stages > n_kchunks + 2) n_stages = n_kchunks + 2;
  if (n_stages < 2) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits: tile does not fit shared memory");
  const unsigned grid = static_cast<unsigned>(n_items < sms ? n_items : sms);
  ProfScope prof(st, prof_kind, (int64_t)B * Q * Ncols);
#define MSDA_TC4_LAUNCH(ATM)                                                                                                            \
  do {                                                                                                                                  \
    if (int rc = ensure_func_attr(mask_fwd_tc4_kernel<OT, kTransB, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) return rc; \
    mask_fwd_tc4_kernel<OT, kTransB, ATM><<<grid, kTc4Threads, 1024 + n_stages * stage + out_bytes, st>>>(                                 \
        map_plane, map_rows, map_out, Q, n_kchunks, QS, QN, n_qchunks, n_tiles_n, static_cast<int>(n_items), n_stages,                  \
        option("mask_debug") != 2);                                                                                                     \
  } while (0)
  if (a_tm) MSDA_TC4_LAUNCH(true);
  else MSDA_TC4_LAUNCH(false);
#undef MSDA_TC4_LAUNCH
  return after_launch("mask_fwd_tc4_kernel");
}

template <typename IT, typename OT>
static int launch_mask_simt(cudaStream_t st, const void* coeff, const void* proto, void* out, int B, int Q, int K,
                            int64_t Ncols) {
  const dim3 grid(static_cast<unsigned>((Ncols + kMaskTN - 1) / kMaskTN), (Q + kMaskTQ - 1) / kMaskTQ, B);
  const int vec_ok = (Ncols % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  ProfScope prof(st, MSDA_PROF_MASK_FWD, (int64_t)B * Q * Ncols);
  mask_fwd_simt_kernel<IT, OT><<<grid, 256, 0, st>>>(static_cast<const IT*>(coeff), static_cast<const IT*>(proto),
                                                      static_cast<OT*>(out), Q, K, Ncols, vec_ok);
  return after_launch("mask_fwd_simt_kernel");
}

// Matcher mask costs on the tensor cores (match_cost_tc.cuh).  One launch: Q <= 256 queries, G <= 16 targets.
size_t match_cost_tc_workspace_floats() { return static_cast<size_t>(kMtMaxCtas + 1) * kMtWsPerCta; }    // per-CTA blocks + their sum
bool match_cost_tc_eligible(const void* coeff, const void* proto, const void* tgt, int K, int64_t Ncols) {
  if (K < 4 || K > 32 || K % 4 != 0) return false;                       // 16-byte global strides, one 32-deep reduction chunk
  if (Ncols % 4 != 0 || Ncols >= (int64_t(1) << 31) - kMtTile) return false;
  return ((reinterpret_cast<uintptr_t>(coeff) | reinterpret_cast<uintptr_t>(proto) | reinterpret_cast<uintptr_t>(tgt)) & 15u) == 0;
}
int match_cost_tc_dispatch(cudaStream_t st, const float* coeff, const float* proto, const float* tgt, int Q, int K, int G, int64_t Ncols,
                           float* ws, float* cost_bce, float* cost_dice, int ld) {
  if (Q < 1 || Q > kMtMaxQ || G < 1 || G > kMtGP) return fail(MSDA_ERR_INVALID_ARG, "match_cost_tc: Q=%d G=%d per launch", Q, G);
  CUtensorMap map_plane, map_coeff;
  if (int rc = make_map_mn_f32(&map_plane, proto, (uint64_t)Ncols, (uint64_t)K, 1)) return rc;
  if (int rc = make_map_in(&map_coeff, coeff, MSDA_F32, (uint64_t)K, (uint64_t)Q, 1, (uint32_t)kMtMaxQ)) return rc;
  const int64_t n_items = (Ncols + kMtTile - 1) / kMtTile;
  int ctas = option("consumer_ctas") > 0 ? option("consumer_ctas") : sm_count();
  if (ctas > kMtMaxCtas) ctas = kMtMaxCtas;
  if (ctas > n_items) ctas = static_cast<int>(n_items);
  {
    ProfScope prof(st, 4, static_cast<int64_t>(Q) * Ncols);
#define MSDA_MT_LAUNCH(GPV)                                                                                                             \
    do {                                                                                                                                \
      if (int rc = ensure_func_attr(match_cost_tc_kernel<GPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMtSmemBytes))) return rc; \
      if (int rc = check_cuda(launch_kernel(match_cost_tc_kernel<GPV>, dim3(ctas), dim3(kMtThreads), kMtSmemBytes, st, map_plane, map_coeff, tgt, Q, G, Ncols, static_cast<int>(n_items), ws), "launch")) return rc; \
    } while (0)
    if (G <= 4) MSDA_MT_LAUNCH(4); else if (G <= 8) MSDA_MT_LAUNCH(8); else if (G <= 12) MSDA_MT_LAUNCH(12); else MSDA_MT_LAUNCH(16);
#undef MSDA_MT_LAUNCH
  }
  if (int rc = after_launch("match_cost_tc_kernel")) return rc;
  float* total = ws + static_cast<size_t>(kMtMaxCtas) * kMtWsPerCta;
  if (int rc = check_cuda(launch_kernel(match_cost_tc_reduce_kernel, dim3((kMtWsPerCta + 127) / 128), dim3(128), 0, st, static_cast<const float*>(ws), ctas, total), "launch")) return rc;
  if (int rc = after_launch("match_cost_tc_reduce_kernel")) return rc;
  if (int rc = check_cuda(launch_kernel(match_cost_tc_finalize_kernel, dim3((Q * G + 127) / 128), dim3(128), 0, st, static_cast<const float*>(total), coeff, Q, K, G, Ncols, ld, cost_bce, cost_dice), "launch")) return rc;
  return after_launch("match_cost_tc_finalize_kernel");
}

int mask_forward_dispatch(cudaStream_t st, int in_dtype, int out_dtype, const void* coeff, const void* proto, int B,
                          int Q, int K, int64_t Ncols, void* out) {
  if (B > 65535) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_forward: B=%d > 65535", B);
  const int variant = option("mask_variant");
  const bool tc_ok = mask_tc_eligible(in_dtype, coeff, proto, Q, K, Ncols);
  if (variant == 2 && !tc_ok && !mask_tc3_eligible(in_dtype, coeff, proto, Q, K, Ncols))
    return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_forward: tcgen05 path needs K %% 8 == 0 with K <= 64 (bf16) / 32 (fp32) and 16-byte aligned rows");
  if (variant != 1 && tc_ok) {
    if (out_dtype == MSDA_F32) return launch_mask_tc2<float>(st, in_dtype, coeff, proto, out, B, Q, K, Ncols);
    if (out_dtype == MSDA_F16) return launch_mask_tc2<__half>(st, in_dtype, coeff, proto, out, B, Q, K, Ncols);
    return launch_mask_tc2<__nv_bfloat16>(st, in_dtype, coeff, proto, out, B, Q, K, Ncols);
  }
  if (variant != 1 && mask_tc3_eligible(in_dtype, coeff, proto, Q, K, Ncols)) {
    if (out_dtype == MSDA_F32) return launch_mask_tc4<float, false>(st, coeff, proto, out, B, Q, K, Ncols);
    return launch_mask_tc4<__nv_bfloat16, false>(st, coeff, proto, out, B, Q, K, Ncols);
  }
  if (in_dtype == MSDA_F32 && out_dtype == MSDA_F32) return launch_mask_simt<float, float>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_F32 && out_dtype == MSDA_BF16) return launch_mask_simt<float, __nv_bfloat16>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_BF16 && out_dtype == MSDA_F32) return launch_mask_simt<__nv_bfloat16, float>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_F16 && out_dtype == MSDA_F32) return launch_mask_simt<__half, float>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_F16) return launch_mask_simt<__half, __half>(st, coeff, proto, out, B, Q, K, Ncols);
  return launch_mask_simt<__nv_bfloat16, __nv_bfloat16>(st, coeff, proto, out, B, Q, K, Ncols);
}

// grad_coeff on the tensor cores (mask_grad_coeff_tc_kernel); grad_coeff must already be zeroed
static int launch_mask_grad_coeff_tc(cudaStream_t st, const void* proto, const void* grad_out, void* grad_coeff, int B,
                                     int Q, int K, int64_t Ncols) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int KP = (K + 15) / 16 * 16;
  const int MH = Q > 128 ? 2 : 1;
  const int n_qblocks = (Q + 128 * MH - 1) / (128 * MH);
  const int n_chunks = static_cast<int>((Ncols + 31) / 32);
  const size_t b_bytes = (static_cast<size_t>(KP) * 128 + 1023) & ~size_t(1023);
  const size_t stage = 2 * static_cast<size_t>(MH) * kGcTcHalfBytes + 2 * b_bytes;
  int n_stages = static_cast<int>((224 * 1024) / stage);
  if (n_stages > kGcTcMaxStages) n_stages = kGcTcMaxStages;
  if (n_stages < 2) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_backward: K=%d too large for the tensor-core kernel", K);
  int64_t slices = sms / ((int64_t)B * n_qblocks);
  if (slices < 1) slices = 1;
  if (slices > n_chunks) slices = n_chunks;
  const int cps = static_cast<int>((n_chunks + slices - 1) / slices);
  slices = (n_chunks + cps - 1) / cps;
  CUtensorMap map_go, map_proto;
  if (int rc = make_map_in(&map_go, grad_out, MSDA_F32, (uint64_t)Ncols, (uint64_t)Q, (uint64_t)B, 128u)) return rc;
  if (int rc = make_map_in(&map_proto, proto, MSDA_F32, (uint64_t)Ncols, (uint64_t)K, (uint64_t)B, (uint32_t)KP)) return rc;
  if (int rc = ensure_func_attr(mask_grad_coeff_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) return rc;
  const dim3 grid(static_cast<unsigned>(slices), static_cast<unsigned>(n_qblocks), static_cast<unsigned>(B));
  long long* dbg = nullptr;
  if (option("mask_debug") == 1) {
    static long long* d_dbg = nullptr;
    if (!d_dbg) cudaMalloc(&d_dbg, 5 * 16 * sizeof(long long));
    cudaMemsetAsync(d_dbg, 0, 5 * 16 * sizeof(long long), st);
    dbg = d_dbg;
    g_mask_dbg = d_dbg;
  }
  mask_grad_coeff_tc_kernel<<<grid, kGcTcThreads, 1024 + n_stages * stage, st>>>(map_go, map_proto, static_cast<float*>(grad_coeff), Q, K, KP, MH,
                                                                                n_stages, n_chunks, cps, option("mask_debug") != 2, dbg);
  return after_launch("mask_grad_coeff_tc_kernel");
}

int mask_backward_dispatch(cudaStream_t st, int dtype, const void* coeff, const void* proto, const void* grad_out,
                           int B, int Q, int K, int64_t Ncols, void* grad_coeff, void* grad_proto) {
  if (dtype != MSDA_F32) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_backward: only MSDA_F32 is implemented");
  if (B > 65535) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_backward: B=%d > 65535", B);
  if (grad_coeff)
    if (int rc = check_cuda(cudaMemsetAsync(grad_coeff, 0, (size_t)B * Q * K * sizeof(float), st), "cudaMemsetAsync(grad_coeff)")) return rc;
  if ((int64_t)B * Q * Ncols == 0) {
    if (grad_proto) return check_cuda(cudaMemsetAsync(grad_proto, 0, (size_t)B * K * Ncols * sizeof(float), st), "cudaMemsetAsync(grad_proto)");
    return 0;
  }
  if (!grad_coeff && !grad_proto) return 0;
  ProfScope prof(st, MSDA_PROF_MASK_BWD, (int64_t)B * Q * Ncols);
  const unsigned kblocks = (K + 31) / 32;
  const bool tc_ok = option("mask_variant") != 1 && K % 4 == 0 && K <= 128 && Ncols % 4 == 0 && Ncols < (int64_t(1) << 31) && Q < 65536 * 128 &&
                     ((reinterpret_cast<uintptr_t>(coeff) | reinterpret_cast<uintptr_t>(proto) | reinterpret_cast<uintptr_t>(grad_out) |
                       reinterpret_cast<uintptr_t>(grad_coeff) | reinterpret_cast<uintptr_t>(grad_proto)) & 15u) == 0;
  if (grad_coeff && tc_ok) {
    if (int rc = launch_mask_grad_coeff_tc(st, proto, grad_out, grad_coeff, B, Q, K, Ncols)) return rc;
  } else if (grad_coeff) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int q_blocks = (Q + kGcQ - 1) / kGcQ;
    int64_t slices = (2LL * sms) / ((int64_t)B * q_blocks * kblocks);
    if (slices < 1) slices = 1;
    int64_t cols = ((Ncols + slices - 1) / slices + kGcNC - 1) / kGcNC * kGcNC;
    slices = (Ncols + cols - 1) / cols;
    const dim3 grid(static_cast<unsigned>(slices), kblocks, static_cast<unsigned>(B * q_blocks));
    mask_grad_coeff_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(proto), static_cast<const float*>(grad_out),
                                                 static_cast<float*>(grad_coeff), Q, K, Ncols, cols, q_blocks);
    if (int rc = after_launch("mask_grad_coeff_kernel")) return rc;
  }
  // grad_proto[b, k, n] = sum_q coeff[b, q, k] * grad_out[b, q, n] on the tensor cores (3xTF32): the forward kernel with
  // rows = k, reduction = q (7 chunks of 32 for Q = 196) and the row operand transposed on chip
  if (grad_proto && tc_ok) {
    return launch_mask_tc4<float, true>(st, coeff, grad_out, grad_proto, B, K, Q, Ncols, -1);
  }
  if (grad_proto) {
    const dim3 grid(static_cast<unsigned>((Ncols + kGpTN - 1) / kGpTN), kblocks, B);
    mask_grad_proto_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(coeff), static_cast<const float*>(grad_out),
                                                 static_cast<float*>(grad_proto), Q, K, Ncols);
    if (int rc = after_launch("mask_grad_proto_kernel")) return rc;
  }
  return 0;
}

// ------------------------------------------------------------------------------ Linear layers (3xTF32 GEMM, gemm3x.cuh)
// fp32 [d1][d0] matrix (d0 contiguous), box {32 columns, 128 rows}, 128B swizzle: output tiles of gemm3x_kernel
static int make_map_c(CUtensorMap* map, void* base, uint64_t d0, uint64_t d1) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {d0, d1, 1};
  const cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
  const cuuint32_t box[3] = {32, kG3Tile, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MSDA_ERR_CUDA, "cuTensorMapEncodeTiled(gemm out) failed (CUresult %d)", static_cast<int>(r));
  return 0;
}

// Per-tile flags of the stream-K form: one zero-initialised ring per device, a fresh stretch of it per launch (launches that overlap
// on different streams or graph branches never share flags; the kernel leaves its flags zero again).  Allocated on first use outside
// of stream capture; until then (first call inside a capture) the GEMMs run in the round-robin tile form.
namespace {
constexpr int kFlagRing = 1 << 20;                     // 4 MB: thousands of launches between two uses of the same flags
struct FlagRing { int* base = nullptr; int cursor = 0; };
std::mutex g_flag_mu;
FlagRing g_flag_ring[64];
}  // namespace

static int* gemm_tile_flags(cudaStream_t st, int tiles) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || tiles > kFlagRing / 4) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(g_flag_mu);
  FlagRing& ring = g_flag_ring[dev];
  if (!ring.base) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    int* p = nullptr;
    if (cudaMalloc(&p, kFlagRing * sizeof(int)) != cudaSuccess || cudaMemset(p, 0, kFlagRing * sizeof(int)) != cudaSuccess) {
      cudaGetLastError();
      if (p) cudaFree(p);
      return nullptr;
    }
    ring.base = p;
  }
  if (ring.cursor + tiles > kFlagRing) ring.cursor = 0;
  int* out = ring.base + ring.cursor;
  ring.cursor += tiles;
  return out;
}

// C[M x N] (+)= A * B.  a_mn: A is stored [K][M] (else [M][K]); b_mn: B is stored [K][N] (else [N][K]).  splits > 1: the reduction
// is cut into `splits` ranges whose partial tiles are added into C (which must be zero) with TMA reduce-add stores; splits == 1:
// the (tile, chunk) units are dealt out as one contiguous range per CTA (stream-K, see gemm3x.cuh).
template <bool kAMn, bool kBMn>
static int launch_gemm3x(cudaStream_t st, const char* who, const void* A, const void* Bm, void* C, int64_t M, int64_t N, int64_t K,
                         const float* bias, const unsigned char* row_mask, int splits, float* col_sum_a = nullptr,
                         G3Packed pk = G3Packed{nullptr, nullptr, nullptr, 0, 0, 0}) {
  const int sms = sm_count();
  const int n_kchunks = static_cast<int>((K + 31) / 32);
  const int tiles_m = static_cast<int>((M + kG3Tile - 1) / kG3Tile), tiles_n = static_cast<int>((N + kG3Tile - 1) / kG3Tile);
  if (splits < 1) splits = 1;
  if (splits > n_kchunks) splits = n_kchunks;
  const int cps = (n_kchunks + splits - 1) / splits;
  splits = (n_kchunks + cps - 1) / cps;
  const int64_t tiles = (int64_t)tiles_m * tiles_n;
  int64_t n_items = tiles * splits;
  int mode = splits > 1 ? kG3ModeReduce : kG3ModeStore;
  int* flags = nullptr;
  // contiguous ranges pay when whole tiles dealt round-robin would leave the SMs idle for >= 4 chunk times on average (or there are
  // fewer tiles than SMs): the flag hand-over of the split tiles costs about one (A/B per shape: profiles/r02ao_linear_bench_*.json)
  const int64_t idle_units = ((tiles + sms - 1) / sms * sms - tiles) * n_kchunks;
  const bool worth = tiles < sms || idle_units >= 4 * (int64_t)sms || option("gemm_stream_k") == 2;
  if (splits == 1 && pk.packed == nullptr && tiles * n_kchunks < (int64_t(1) << 31) && option("gemm_stream_k") != 0 && worth &&   // packed epilogue: bf16 rounding needs whole tiles
      (flags = gemm_tile_flags(st, static_cast<int>(tiles))) != nullptr) {
    mode = kG3ModeStreamK;
    n_items = tiles * n_kchunks;
  }
  if (n_items >= (int64_t(1) << 31)) return fail(MSDA_ERR_UNSUPPORTED, "%s: too many tiles", who);
  CUtensorMap map_a, map_b, map_c;
  if (kAMn) { if (int rc = make_map_mn_f32(&map_a, A, (uint64_t)M, (uint64_t)K, 1)) return rc; }
  else { if (int rc = make_map_in(&map_a, A, MSDA_F32, (uint64_t)K, (uint64_t)M, 1, kG3Tile)) return rc; }
  if (kBMn) { if (int rc = make_map_mn_f32(&map_b, Bm, (uint64_t)N, (uint64_t)K, 1)) return rc; }
  else { if (int rc = make_map_in(&map_b, Bm, MSDA_F32, (uint64_t)K, (uint64_t)N, 1, kG3Tile)) return rc; }
  if (pk.packed != nullptr) map_c = map_a;                  // never stored through: the epilogue writes the packed layout itself
  else if (int rc = make_map_c(&map_c, C, (uint64_t)N, (uint64_t)M)) return rc;
  if (int rc = ensure_func_attr(gemm3x_kernel<kAMn, kBMn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kG3SmemBytes)) return rc;
  const unsigned grid = static_cast<unsigned>(n_items < sms ? n_items : sms);
  if (int rc = check_cuda(launch_kernel(gemm3x_kernel<kAMn, kBMn>, dim3(grid), dim3(kG3Threads), kG3SmemBytes, st, map_a, map_b, map_c, bias,
                                        row_mask, col_sum_a, flags, (int)M, (int)N, n_kchunks, cps, tiles_m, tiles_n, (int)n_items, mode, pk),
                          "gemm3x_kernel launch")) return rc;
  return after_launch("gemm3x_kernel");
}

// aligned16(): msda_launch.cuh

int64_t gemm_flag_timeouts() {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_g3_flag_timeouts, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
  return static_cast<int64_t>(v);
}

int linear_forward_dispatch(cudaStream_t st, const void* x, const void* w, const void* bias, const unsigned char* row_mask,
                            int64_t rows, int in_f, int out_f, void* y) {
  if (rows >= (int64_t(1) << 31) - kG3Tile) return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_forward: rows=%lld too large", (long long)rows);
  if (in_f % 4 != 0 || out_f % 4 != 0 || !aligned16(x) || !aligned16(w) || !aligned16(y))
    return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_forward: in_features / out_features must be multiples of 4 and the tensors 16-byte aligned");
  if (rows == 0) return 0;
  return launch_gemm3x<false, false>(st, "tc_linear_forward", x, w, y, rows, out_f, in_f, static_cast<const float*>(bias), row_mask, 1);
}

int linear_forward_packed_dispatch(cudaStream_t st, const void* x, const void* w, const void* bias, const unsigned char* row_mask, int N,
                                   int S, int in_f, int heads, const int64_t* shapes, const int64_t* level_start, int L, void* packed) {
  const int64_t rows = (int64_t)N * S;
  if (rows >= (int64_t(1) << 31) - kG3Tile || (int64_t)N * 2 * S * heads >= (int64_t(1) << 32))
    return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_forward_packed: N*S=%lld too large", (long long)rows);
  if (in_f % 4 != 0 || !aligned16(x) || !aligned16(w) || !aligned16(packed) || L > kMaxLevels)
    return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_forward_packed: in_features must be a multiple of 4, L <= %d and the tensors 16-byte aligned", kMaxLevels);
  if (rows == 0) return 0;
  const G3Packed pk{static_cast<uint4*>(packed), shapes, level_start, L, S, heads};
  return launch_gemm3x<false, false>(st, "tc_linear_forward_packed", x, w, nullptr, rows, (int64_t)heads * 32, in_f, static_cast<const float*>(bias), row_mask, 1,
                                     nullptr, pk);
}

int linear_backward_dispatch(cudaStream_t st, const void* gy, const void* x, const void* w, int64_t rows, int in_f, int out_f,
                             void* gx, void* gw, void* gb, bool* gb_done) {
  if (gb_done) *gb_done = false;
  if (rows >= (int64_t(1) << 31) - kG3Tile) return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_backward: rows=%lld too large", (long long)rows);
  if (in_f % 4 != 0 || out_f % 4 != 0 || !aligned16(gy) || (gx && (!aligned16(gx) || !aligned16(w))) || (gw && (!aligned16(gw) || !aligned16(x))))
    return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_backward: in_features / out_features must be multiples of 4 and the tensors 16-byte aligned");
  const bool fuse_bias = gb != nullptr && gw != nullptr;                    // the weight-gradient GEMM stages grad_y: its column sums come for free
  if (gw)
    if (int rc = check_cuda(cudaMemsetAsync(gw, 0, (size_t)out_f * in_f * sizeof(float), st), "cudaMemsetAsync(grad_weight)")) return rc;
  if (fuse_bias) {
    if (int rc = check_cuda(cudaMemsetAsync(gb, 0, (size_t)out_f * sizeof(float), st), "cudaMemsetAsync(grad_bias)")) return rc;
    if (gb_done) *gb_done = true;
  }
  if (rows == 0) return 0;
  if (gx) {                                                 // dx[r, i] = sum_o dy[r, o] W[o, i]
    if (int rc = launch_gemm3x<false, true>(st, "tc_linear_backward", gy, w, gx, rows, in_f, out_f, nullptr, nullptr, 1)) return rc;
  }
  if (gw) {                                                 // dW[o, i] = sum_r dy[r, o] x[r, i]: reduction over the rows, split across the SMs
    const int tiles = ((out_f + kG3Tile - 1) / kG3Tile) * ((in_f + kG3Tile - 1) / kG3Tile);
    const int splits = std::max(2, sm_count() / tiles);     // at most one (tile, reduction range) item per SM: 6 tiles x 24, not 25
    if (int rc = launch_gemm3x<true, true>(st, "tc_linear_backward", gy, x, gw, out_f, in_f, rows, nullptr, nullptr, splits,
                                           fuse_bias ? static_cast<float*>(gb) : nullptr)) return rc;
  }
  return 0;
}

}  // namespace msda
