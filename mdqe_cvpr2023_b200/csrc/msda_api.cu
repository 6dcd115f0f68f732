// C ABI of libmsda_b200.so: argument validation, kernel selection and launch for the multi-scale
// deformable attention operator, plus the host-buffer entry points.  See include/msda_b200.h for the
// contract and the reference interfaces each entry replaces.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <type_traits>
#include <vector>

#include "msda_launch.cuh"

namespace msda {

// ------------------------------------------------------------------------------------ bookkeeping
static thread_local char t_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return 0;
  return fail(MSDA_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
}

int after_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), kernel_name);
}

static Options g_opt;
const Options& options() { return g_opt; }

static std::atomic<int>* find_option(const char* key) {
  if (!key) return nullptr;
  if (!strcmp(key, "fwd_variant")) return &g_opt.fwd_variant;
  if (!strcmp(key, "bwd_variant")) return &g_opt.bwd_variant;
  if (!strcmp(key, "chunk_pairs")) return &g_opt.chunk_pairs;
  if (!strcmp(key, "mask_variant")) return &g_opt.mask_variant;
  if (!strcmp(key, "profile")) return &g_opt.profile;
  if (!strcmp(key, "mask_debug")) return &g_opt.mask_debug;
  if (!strcmp(key, "host_async")) return &g_opt.host_async;
  if (!strcmp(key, "consumer_ctas")) return &g_opt.consumer_ctas;
  if (!strcmp(key, "consumer_tc")) return &g_opt.consumer_tc;
  if (!strcmp(key, "bwd_merge")) return &g_opt.bwd_merge;
  if (!strcmp(key, "pair_map")) return &g_opt.pair_map;
  if (!strcmp(key, "pdl")) return &g_opt.pdl;
  if (!strcmp(key, "gemm_stream_k")) return &g_opt.gemm_stream_k;
  if (!strcmp(key, "mask_a_tmem")) return &g_opt.mask_a_tmem;
  return nullptr;
}

namespace {
struct AttrKey { const void* func; int attr, device, value; };
std::mutex g_attr_mu;
std::vector<AttrKey> g_attr_done;
}  // namespace

int ensure_func_attr_impl(const void* func, cudaFuncAttribute attr, int value) {
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  std::lock_guard<std::mutex> lock(g_attr_mu);
  for (const AttrKey& k : g_attr_done)
    if (k.func == func && k.attr == static_cast<int>(attr) && k.device == dev && k.value == value) return 0;
  if (int rc = check_cuda(cudaFuncSetAttribute(func, attr, value), "cudaFuncSetAttribute")) return rc;
  g_attr_done.push_back(AttrKey{func, static_cast<int>(attr), dev, value});
  return 0;
}

int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64 && cache[dev].load(std::memory_order_relaxed) > 0) return cache[dev].load(std::memory_order_relaxed);
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); sms = 148; }
  if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

int option(const char* key) {
  std::atomic<int>* o = find_option(key);
  return o ? o->load() : 0;
}

// ---------------------------------------------------------------------------------------- profiling
// Optional per-launch timing with CUDA events recorded on the launching stream, right around the
// kernel (not around memsets or conversions).  Used by bench.py for the roofline figure.
struct ProfRec { cudaEvent_t a, b; int kind; int64_t units; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(cudaStream_t st, int kind, int64_t units) : st_(st), kind_(kind), units_(units) {
  if (kind < 0 || !g_opt.profile.load()) return;      // kind < 0: the caller already times this launch
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  if (cudaEventCreate(&a_) != cudaSuccess || cudaEventCreate(&b_) != cudaSuccess) { a_ = b_ = nullptr; return; }
  cudaEventRecord(a_, st_);
  on_ = true;
}
ProfScope::~ProfScope() {
  if (!on_) return;
  cudaEventRecord(b_, st_);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_prof.size() < (1u << 16)) g_prof.push_back(ProfRec{a_, b_, kind_, units_});
  else { cudaEventDestroy(a_); cudaEventDestroy(b_); }
}

// -------------------------------------------------------------------------------------- validation

static int validate(const char* who, int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                    const void* loc, const void* aw, int N, int S, int M, int D, int L, int Lq, int P) {
  if (dtype_size(dtype) == 0) return fail(MSDA_ERR_INVALID_ARG, "%s: unknown dtype %d", who, dtype);
  if (N < 0 || S < 0 || M <= 0 || D <= 0 || L <= 0 || Lq < 0 || P <= 0)
    return fail(MSDA_ERR_INVALID_ARG, "%s: bad sizes N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d", who, N, S, M, D, L, Lq, P);
  if (!shapes || !lsi) return fail(MSDA_ERR_INVALID_ARG, "%s: spatial_shapes / level_start_index is NULL", who);
  if ((int64_t)N * S * M * D > 0 && !value) return fail(MSDA_ERR_INVALID_ARG, "%s: value is NULL", who);
  if ((int64_t)N * Lq > 0 && (!loc || !aw)) return fail(MSDA_ERR_INVALID_ARG, "%s: sampling_loc / attn_weight is NULL", who);
  return 0;
}

// instantiated in msda_launch_*.cu
MSDA_LAUNCH_EXTERN(extern, float, float, float)
MSDA_LAUNCH_EXTERN(extern, __nv_bfloat16, __nv_bfloat16, float)
MSDA_LAUNCH_EXTERN(extern, __nv_bfloat16, float, float)
MSDA_LAUNCH_EXTERN(extern, double, double, double)

}  // namespace msda

using namespace msda;

// ===================================================================================== public ABI
// column sums of grad_y: thread = (4 columns, one of 4 row phases); rows are split across CTAs, partial sums meet in grad_bias
__global__ void __launch_bounds__(256) linear_bias_grad_kernel(const float* __restrict__ gy, int64_t rows, int out_f, int64_t rows_per_cta,
                                                                float* __restrict__ gb) {
  __shared__ float4 s_part[4][64];
  const int c4 = blockIdx.y * 64 + (threadIdx.x & 63), phase = threadIdx.x >> 6;
  const bool col_ok = c4 * 4 < out_f;
  const int64_t r_begin = blockIdx.x * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col_ok) {
    const float4* base = reinterpret_cast<const float4*>(gy) + c4;
    const int64_t stride4 = out_f / 4;
    for (int64_t r = r_begin + phase; r < r_end; r += 16) {             // four independent loads in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + 4 * u;
        if (rr < r_end) {
          const float4 v = __ldg(base + rr * stride4);
          acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
        }
      }
    }
  }
  float4 a = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                         (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  s_part[phase][threadIdx.x & 63] = a;
  __syncthreads();
  if (phase == 0 && col_ok) {
    const int c = threadIdx.x & 63;
    const float4 b1 = s_part[1][c], b2 = s_part[2][c], b3 = s_part[3][c];
    float* dst = gb + 4 * c4;
    atomicAdd(dst + 0, (a.x + b1.x) + (b2.x + b3.x));
    atomicAdd(dst + 1, (a.y + b1.y) + (b2.y + b3.y));
    atomicAdd(dst + 2, (a.z + b1.z) + (b2.z + b3.z));
    atomicAdd(dst + 3, (a.w + b1.w) + (b2.w + b3.w));
  }
}

extern "C" {

int msda_abi_version(void) { return MSDA_B200_ABI_VERSION; }
#ifdef MSDA_DBG_MASK
int msda_debug_set_mask(int mask) { return check_cuda(cudaMemcpyToSymbol(g_dbg_mask, &mask, sizeof(int)), "msda_debug_set_mask"); }
#endif
const char* msda_last_error(void) { return t_err; }

int msda_set_option(const char* key, int value) {
  std::atomic<int>* o = find_option(key);
  if (!o) return fail(MSDA_ERR_INVALID_ARG, "msda_set_option: unknown key '%s'", key ? key : "(null)");
  o->store(value);
  return 0;
}

int msda_get_option(const char* key, int* value) {
  std::atomic<int>* o = find_option(key);
  if (!o || !value) return fail(MSDA_ERR_INVALID_ARG, "msda_get_option: unknown key '%s'", key ? key : "(null)");
  *value = o->load();
  return 0;
}

int msda_profile_read(int kind, int64_t min_units, double* total_ms, int64_t* count) {
  if (!total_ms || !count) return fail(MSDA_ERR_INVALID_ARG, "msda_profile_read: NULL output");
  std::lock_guard<std::mutex> lock(g_prof_mu);
  double tot = 0.0;
  int64_t n = 0;
  std::vector<ProfRec> keep;
  for (const ProfRec& r : g_prof) {
    if (r.kind != kind) { keep.push_back(r); continue; }
    if (int rc = check_cuda(cudaEventSynchronize(r.b), "cudaEventSynchronize(profile)")) return rc;
    float ms = 0.f;
    if (r.units >= min_units && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot += ms; ++n; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.swap(keep);
  *total_ms = tot;
  *count = n;
  return 0;
}

int msda_debug_read(long long* host80) { return mask_debug_copy(host80); }

int64_t msda_launch_count(void) { return g_launches.load(); }
void msda_launch_count_reset(void) { g_launches.store(0); }
int64_t msda_gemm_flag_timeouts(void) { return gemm_flag_timeouts(); }

static int grouped_supported(const char* who, int dtype, const Problem& pb, bool fast) {
  if (pb.G == 1 && pb.scale == 1.f) return 0;            // the plain operator: every kernel family implements it
  if (pb.G < 1 || pb.G * pb.L > kMaxLevels) return fail(MSDA_ERR_UNSUPPORTED, "%s: G*L = %d exceeds %d level tables", who, pb.G * pb.L, kMaxLevels);
  if (!fast || !fast2_lp(pb.L * pb.P) || dtype == MSDA_F64)
    return fail(MSDA_ERR_UNSUPPORTED, "%s: the grouped form needs D in {32,24}, L*P in {8,12,16}, fp32/bf16 and 16-byte aligned tensors", who);
  return 0;
}

static int fused_supported(const char* who, int dtype, const Problem& pb, bool fast) {
  if (pb.fz.ref == nullptr) return 0;
  const int lp = pb.L * pb.P;
  if (!fast || dtype != MSDA_F32 || (lp != 8 && lp != 16))
    return fail(MSDA_ERR_UNSUPPORTED, "%s: the fused prologue needs fp32, D in {32,24}, L*P in {8,16} and 16-byte aligned tensors", who);
  if ((pb.fz.R != 2 && pb.fz.R != 4) || (pb.fz.mode != 0 && pb.fz.mode != 1) || (pb.fz.mode == 1 && (pb.fz.R != 4 || !pb.fz.grid)) || !(pb.fz.scale > 0.f))
    return fail(MSDA_ERR_INVALID_ARG, "%s: bad fused arguments (R=%d mode=%d scale=%g)", who, pb.fz.R, pb.fz.mode, (double)pb.fz.scale);
  return 0;
}

static int forward_impl(const char* who, void* stream, int dtype, const void* value, const int64_t* shapes,
                        const int64_t* level_start, const void* loc, const void* aw, int N, int S, int M, int D, int G, int L,
                        int Lq, int P, float scale, void* out, FusedArgs fz = FusedArgs{nullptr, nullptr, 0, 0, 1.f}) {
  if (int rc = validate(who, dtype, value, shapes, level_start, loc, aw, N, S, M, D, L, Lq, P)) return rc;
  Problem pb{N, S, M, D, L, Lq, P, (int64_t)N * Lq * M, G, scale, fz};
  if (pb.n_pairs == 0) return 0;
  if (!out) return fail(MSDA_ERR_INVALID_ARG, "%s: out is NULL", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fast = (G > 1 || fz.ref || g_opt.fwd_variant.load() != 1) && fast_eligible(dtype, pb, value, out, loc);
  if (int rc = grouped_supported(who, dtype, pb, fast)) return rc;
  if (int rc = fused_supported(who, dtype, pb, fast)) return rc;
  switch (dtype) {
    case MSDA_F32: return launch_fwd<float, float>(st, pb, fast, value, shapes, level_start, loc, aw, out);
    case MSDA_BF16: return launch_fwd<__nv_bfloat16, __nv_bfloat16>(st, pb, fast, value, shapes, level_start, loc, aw, out);
    case MSDA_BF16_LOC32: return launch_fwd<__nv_bfloat16, float>(st, pb, fast, value, shapes, level_start, loc, aw, out);
    case MSDA_F64: return launch_fwd<double, double>(st, pb, false, value, shapes, level_start, loc, aw, out);
  }
  return fail(MSDA_ERR_INVALID_ARG, "%s: unknown dtype %d", who, dtype);
}

int msda_forward(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                 const void* loc, const void* aw, int N, int S, int M, int D, int L, int Lq, int P, void* out) {
  return forward_impl("msda_forward", stream, dtype, value, shapes, level_start, loc, aw, N, S, M, D, 1, L, Lq, P, 1.f, out);
}

int msda_forward_grouped(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                         const void* loc, const void* aw, int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                         void* out) {
  return forward_impl("msda_forward_grouped", stream, dtype, value, shapes, level_start, loc, aw, N, S, M, D, G, L, Lq, P, scale, out);
}

size_t msda_backward_workspace_bytes(int dtype, int N, int S, int M, int D) {
  if (dtype == MSDA_BF16 || dtype == MSDA_BF16_LOC32) return (size_t)N * S * M * D * sizeof(float);
  return 0;
}

static int backward_impl(const char* who, void* stream, int dtype, const void* value, const int64_t* shapes,
                         const int64_t* level_start, const void* loc, const void* aw, const void* grad_out, int N, int S, int M,
                         int D, int G, int L, int Lq, int P, float scale, void* grad_value, void* grad_loc, void* grad_aw,
                         void* workspace, size_t workspace_bytes, FusedArgs fz = FusedArgs{nullptr, nullptr, 0, 0, 1.f}, int flags = 0) {
  if (int rc = validate(who, dtype, value, shapes, level_start, loc, aw, N, S, M, D, L, Lq, P)) return rc;
  Problem pb{N, S, M, D, L, Lq, P, (int64_t)N * Lq * M, G, scale, fz};
  const size_t n_value = (size_t)N * S * M * D;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_value > 0 && !grad_value) return fail(MSDA_ERR_INVALID_ARG, "%s: grad_value is NULL", who);
  const bool is_bf16 = dtype == MSDA_BF16 || dtype == MSDA_BF16_LOC32;
  const size_t need = msda_backward_workspace_bytes(dtype, N, S, M, D);
  if (need > 0 && (!workspace || workspace_bytes < need))
    return fail(MSDA_ERR_WORKSPACE, "%s: bf16 needs a %zu-byte fp32 workspace (got %zu)", who, need, workspace_bytes);
  void* acc = is_bf16 ? workspace : grad_value;
  const size_t acc_bytes = is_bf16 ? need : n_value * dtype_size(dtype);
  if (flags & ~MSDA_BWD_ACC_ZEROED) return fail(MSDA_ERR_INVALID_ARG, "%s: unknown flags 0x%x", who, flags);
  if (acc_bytes > 0 && !(flags & MSDA_BWD_ACC_ZEROED))    // MSDA_BWD_ACC_ZEROED: the caller zero-filled it off the critical path
    if (int rc = check_cuda(cudaMemsetAsync(acc, 0, acc_bytes, st), "cudaMemsetAsync(grad_value)")) return rc;
  if (pb.n_pairs == 0) {
    if (is_bf16 && n_value > 0)
      return check_cuda(cudaMemsetAsync(grad_value, 0, n_value * 2, st), "cudaMemsetAsync(grad_value)");
    return 0;
  }
  if (!grad_out || !grad_loc || !grad_aw) return fail(MSDA_ERR_INVALID_ARG, "%s: grad_out / grad_loc / grad_aw is NULL", who);
  if (n_value == 0) {
    // nothing to sample from: every corner is out of range, all gradients are zero (the fast backward gathers unconditionally
    // from row 0 for such corners, so it must never see an empty value tensor)
    const size_t n_smp = (size_t)pb.n_pairs * L * P, ls = loc_dtype_size(dtype);
    if (int rc = check_cuda(cudaMemsetAsync(grad_loc, 0, n_smp * 2 * ls, st), "cudaMemsetAsync(grad_loc)")) return rc;
    return check_cuda(cudaMemsetAsync(grad_aw, 0, n_smp * ls, st), "cudaMemsetAsync(grad_aw)");
  }
  const bool fast = (G > 1 || fz.ref || g_opt.bwd_variant.load() != 1) && fast_eligible(dtype, pb, value, grad_out, acc) &&
                    aligned16(loc) && aligned16(grad_loc);
  if (int rcg = grouped_supported(who, dtype, pb, fast)) return rcg;
  if (int rcf = fused_supported(who, dtype, pb, fast)) return rcf;
  int rc = 0;
  switch (dtype) {
    case MSDA_F32:
      rc = launch_bwd<float, float, float>(st, pb, fast, value, shapes, level_start, loc, aw, grad_out,
                                           static_cast<float*>(acc), grad_loc, grad_aw);
      break;
    case MSDA_BF16:
      rc = launch_bwd<__nv_bfloat16, __nv_bfloat16, float>(st, pb, fast, value, shapes, level_start, loc, aw, grad_out,
                                                           static_cast<float*>(acc), grad_loc, grad_aw);
      break;
    case MSDA_BF16_LOC32:
      rc = launch_bwd<__nv_bfloat16, float, float>(st, pb, fast, value, shapes, level_start, loc, aw, grad_out,
                                                   static_cast<float*>(acc), grad_loc, grad_aw);
      break;
    case MSDA_F64:
      rc = launch_bwd<double, double, double>(st, pb, false, value, shapes, level_start, loc, aw, grad_out,
                                              static_cast<double*>(acc), grad_loc, grad_aw);
      break;
    default:
      return fail(MSDA_ERR_INVALID_ARG, "%s: unknown dtype %d", who, dtype);
  }
  if (rc) return rc;
  if (is_bf16) {
    if (n_value % 4 != 0 || !aligned16(workspace) || (reinterpret_cast<uintptr_t>(grad_value) & 7u))
      return fail(MSDA_ERR_UNSUPPORTED, "%s: bf16 grad_value conversion needs N*S*M*D %% 4 == 0 and aligned buffers", who);
    const int64_t n4 = (int64_t)(n_value / 4);
    const unsigned grid = static_cast<unsigned>((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    cvt_f32_to_bf16_kernel<<<grid, 256, 0, st>>>(static_cast<const float4*>(workspace), static_cast<uint2*>(grad_value), n4);
    return after_launch("cvt_f32_to_bf16_kernel");
  }
  return 0;
}

int msda_backward(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                  const void* loc, const void* aw, const void* grad_out, int N, int S, int M, int D, int L, int Lq,
                  int P, void* grad_value, void* grad_loc, void* grad_aw, void* workspace, size_t workspace_bytes) {
  return backward_impl("msda_backward", stream, dtype, value, shapes, level_start, loc, aw, grad_out, N, S, M, D, 1, L, Lq, P,
                       1.f, grad_value, grad_loc, grad_aw, workspace, workspace_bytes);
}

int msda_backward_grouped(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                          const void* loc, const void* aw, const void* grad_out, int N, int S, int M, int D, int G, int L,
                          int Lq, int P, float scale, void* grad_value, void* grad_loc, void* grad_aw, void* workspace,
                          size_t workspace_bytes) {
  return backward_impl("msda_backward_grouped", stream, dtype, value, shapes, level_start, loc, aw, grad_out, N, S, M, D, G, L,
                       Lq, P, scale, grad_value, grad_loc, grad_aw, workspace, workspace_bytes);
}

int msda_backward_grouped_flags(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                                const void* loc, const void* aw, const void* grad_out, int N, int S, int M, int D, int G, int L,
                                int Lq, int P, float scale, void* grad_value, void* grad_loc, void* grad_aw, void* workspace,
                                size_t workspace_bytes, int flags) {
  return backward_impl("msda_backward_grouped_flags", stream, dtype, value, shapes, level_start, loc, aw, grad_out, N, S, M, D, G, L,
                       Lq, P, scale, grad_value, grad_loc, grad_aw, workspace, workspace_bytes, FusedArgs{nullptr, nullptr, 0, 0, 1.f}, flags);
}

int msda_zero_fill(void* stream, void* ptr, size_t bytes) {
  if (bytes == 0) return 0;
  if (!ptr) return fail(MSDA_ERR_INVALID_ARG, "msda_zero_fill: NULL pointer");
  return check_cuda(cudaMemsetAsync(ptr, 0, bytes, static_cast<cudaStream_t>(stream)), "cudaMemsetAsync(msda_zero_fill)");
}

int msda_fused_forward(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                       const void* ref_points, int R, const void* offsets, const void* logits, const void* grid, int mode,
                       float offset_scale, int N, int S, int M, int D, int G, int L, int Lq, int P, float scale, void* out) {
  if (!ref_points) return fail(MSDA_ERR_INVALID_ARG, "msda_fused_forward: reference points are NULL");
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid), R, mode, offset_scale};
  return forward_impl("msda_fused_forward", stream, dtype, value, shapes, level_start, offsets, logits, N, S, M, D, G, L, Lq, P,
                      scale, out, fz);
}

int msda_fused_backward(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                        const void* ref_points, int R, const void* offsets, const void* logits, const void* grid, int mode,
                        float offset_scale, const void* grad_out, int N, int S, int M, int D, int G, int L, int Lq, int P,
                        float scale, void* grad_value, void* grad_offsets, void* grad_logits) {
  if (!ref_points) return fail(MSDA_ERR_INVALID_ARG, "msda_fused_backward: reference points are NULL");
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid), R, mode, offset_scale};
  return backward_impl("msda_fused_backward", stream, dtype, value, shapes, level_start, offsets, logits, grad_out, N, S, M, D, G,
                       L, Lq, P, scale, grad_value, grad_offsets, grad_logits, nullptr, 0, fz);
}

int msda_fused_backward_flags(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                        const void* ref_points, int R, const void* offsets, const void* logits, const void* grid, int mode,
                        float offset_scale, const void* grad_out, int N, int S, int M, int D, int G, int L, int Lq, int P,
                        float scale, void* grad_value, void* grad_offsets, void* grad_logits, int flags) {
  if (!ref_points) return fail(MSDA_ERR_INVALID_ARG, "msda_fused_backward_flags: reference points are NULL");
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid), R, mode, offset_scale};
  return backward_impl("msda_fused_backward_flags", stream, dtype, value, shapes, level_start, offsets, logits, grad_out, N, S, M, D, G,
                       L, Lq, P, scale, grad_value, grad_offsets, grad_logits, nullptr, 0, fz, flags);
}

static int joint_args(const char* who, const void* qproj, int row_stride, int M, int L, int P) {
  const int64_t lp = (int64_t)M * L * P;
  if (!qproj) return fail(MSDA_ERR_INVALID_ARG, "%s: qproj is NULL", who);
  if (row_stride < 3 * lp || (row_stride & 3) || (lp & 1) || (reinterpret_cast<uintptr_t>(qproj) & 15u))
    return fail(MSDA_ERR_INVALID_ARG, "%s: row_stride=%d must be a multiple of 4 and >= 3*M*L*P=%lld, qproj 16-byte aligned", who, row_stride,
                (long long)(3 * lp));
  return 0;
}

int msda_fused_forward_joint(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                             const void* ref_points, int R, const void* qproj, int row_stride, const void* grid, int mode,
                             float offset_scale, int N, int S, int M, int D, int G, int L, int Lq, int P, float scale, void* out) {
  if (!ref_points) return fail(MSDA_ERR_INVALID_ARG, "msda_fused_forward_joint: reference points are NULL");
  if (int rc = joint_args("msda_fused_forward_joint", qproj, row_stride, M, L, P)) return rc;
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid), R, mode, offset_scale, row_stride};
  const float* q = static_cast<const float*>(qproj);
  return forward_impl("msda_fused_forward_joint", stream, dtype, value, shapes, level_start, q, q + (int64_t)2 * M * L * P, N, S, M, D, G, L,
                      Lq, P, scale, out, fz);
}

int msda_fused_backward_joint(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                              const void* ref_points, int R, const void* qproj, int row_stride, const void* grid, int mode,
                              float offset_scale, const void* grad_out, int N, int S, int M, int D, int G, int L, int Lq, int P,
                              float scale, void* grad_value, void* grad_qproj, int flags) {
  if (!ref_points) return fail(MSDA_ERR_INVALID_ARG, "msda_fused_backward_joint: reference points are NULL");
  if (int rc = joint_args("msda_fused_backward_joint", qproj, row_stride, M, L, P)) return rc;
  if (!grad_qproj || (reinterpret_cast<uintptr_t>(grad_qproj) & 15u))
    return fail(MSDA_ERR_INVALID_ARG, "msda_fused_backward_joint: grad_qproj must be a 16-byte aligned [N*Lq, row_stride] buffer");
  const FusedArgs fz{static_cast<const float*>(ref_points), static_cast<const float*>(grid), R, mode, offset_scale, row_stride};
  const float* q = static_cast<const float*>(qproj);
  float* gq = static_cast<float*>(grad_qproj);
  const int64_t lp2 = (int64_t)2 * M * L * P;
  return backward_impl("msda_fused_backward_joint", stream, dtype, value, shapes, level_start, q, q + lp2, grad_out, N, S, M, D, G, L, Lq, P,
                       scale, grad_value, gq, gq + lp2, nullptr, 0, fz, flags);
}

int mask_logits_forward(void* stream, int in_dtype, int out_dtype, const void* coeff, const void* proto, int B, int Q,
                        int K, int64_t Ncols, void* out) {
  if (B < 0 || Q < 0 || K <= 0 || Ncols < 0) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward: bad sizes B=%d Q=%d K=%d Ncols=%lld", B, Q, K, (long long)Ncols);
  const bool in_ok = in_dtype == MSDA_F32 || in_dtype == MSDA_BF16 || in_dtype == MSDA_F16;
  const bool out_ok = out_dtype == MSDA_F32 || (in_dtype == MSDA_F32 ? out_dtype == MSDA_BF16 : out_dtype == in_dtype);
  if (!in_ok || !out_ok)
    return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward: in_dtype must be MSDA_F32 / MSDA_BF16 / MSDA_F16 and out_dtype MSDA_F32 or "
                "the inputs' 16-bit type (fp32 inputs: MSDA_F32 or MSDA_BF16); got %d -> %d", in_dtype, out_dtype);
  if ((int64_t)B * Q * Ncols == 0) return 0;
  if (!coeff || !proto || !out) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward: NULL tensor");
  return mask_forward_dispatch(static_cast<cudaStream_t>(stream), in_dtype, out_dtype, coeff, proto, B, Q, K, Ncols, out);
}

int mask_logits_backward(void* stream, int dtype, const void* coeff, const void* proto, const void* grad_out, int B,
                         int Q, int K, int64_t Ncols, void* grad_coeff, void* grad_proto) {
  if (B < 0 || Q < 0 || K <= 0 || Ncols < 0) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_backward: bad sizes");
  if (dtype != MSDA_F32 && dtype != MSDA_BF16) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_backward: dtype must be MSDA_F32 or MSDA_BF16");
  if (!coeff || !proto || !grad_out) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_backward: NULL tensor");
  return mask_backward_dispatch(static_cast<cudaStream_t>(stream), dtype, coeff, proto, grad_out, B, Q, K, Ncols, grad_coeff, grad_proto);
}

int tc_linear_forward(void* stream, const void* x, const void* weight, const void* bias, const unsigned char* row_mask,
                      int64_t rows, int in_features, int out_features, void* y) {
  if (rows < 0 || in_features <= 0 || out_features <= 0) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_forward: bad sizes");
  if (rows > 0 && (!x || !weight || !y)) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_forward: NULL tensor");
  return linear_forward_dispatch(static_cast<cudaStream_t>(stream), x, weight, bias, row_mask, rows, in_features, out_features, y);
}

int tc_linear_forward_packed(void* stream, const void* x, const void* weight, const void* bias, const unsigned char* row_mask,
                             int N, int S, int in_features, int heads, const int64_t* shapes, const int64_t* level_start, int L,
                             void* packed) {
  if (N < 0 || S < 0 || in_features <= 0 || heads <= 0 || L <= 0) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_forward_packed: bad sizes");
  if (!shapes || !level_start || ((int64_t)N * S > 0 && (!x || !weight || !packed))) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_forward_packed: NULL tensor");
  return linear_forward_packed_dispatch(static_cast<cudaStream_t>(stream), x, weight, bias, row_mask, N, S, in_features, heads, shapes,
                                        level_start, L, packed);
}

int tc_linear_bias_grad(void* stream, const void* grad_y, int64_t rows, int out_features, void* grad_bias) {
  if (rows < 0 || out_features <= 0 || !grad_bias || (rows > 0 && !grad_y)) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_bias_grad: bad arguments");
  if (out_features % 4 != 0 || (reinterpret_cast<uintptr_t>(grad_y) & 15u))
    return fail(MSDA_ERR_UNSUPPORTED, "tc_linear_bias_grad: out_features must be a multiple of 4 and grad_y 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemsetAsync(grad_bias, 0, static_cast<size_t>(out_features) * sizeof(float), st), "cudaMemsetAsync(grad_bias)")) return rc;
  if (rows == 0) return 0;
  const int col_blocks = (out_features / 4 + 63) / 64;
  int64_t ctas = (2LL * sm_count() + col_blocks - 1) / col_blocks;        // ~2 CTAs per SM in total
  if (ctas > (rows + 15) / 16) ctas = (rows + 15) / 16;
  if (ctas < 1) ctas = 1;
  const int64_t rows_per_cta = (rows + ctas - 1) / ctas;
  linear_bias_grad_kernel<<<dim3(static_cast<unsigned>(ctas), col_blocks), 256, 0, st>>>(static_cast<const float*>(grad_y), rows, out_features,
                                                                                         rows_per_cta, static_cast<float*>(grad_bias));
  return after_launch("linear_bias_grad_kernel");
}

int tc_linear_backward(void* stream, const void* grad_y, const void* x, const void* weight, int64_t rows, int in_features,
                       int out_features, void* grad_x, void* grad_weight) {
  if (rows < 0 || in_features <= 0 || out_features <= 0) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_backward: bad sizes");
  if (rows > 0 && (!grad_y || (grad_x && !weight) || (grad_weight && !x))) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_backward: NULL tensor");
  return linear_backward_dispatch(static_cast<cudaStream_t>(stream), grad_y, x, weight, rows, in_features, out_features, grad_x, grad_weight);
}

int tc_linear_backward_bias(void* stream, const void* grad_y, const void* x, const void* weight, int64_t rows, int in_features,
                            int out_features, void* grad_x, void* grad_weight, void* grad_bias) {
  if (rows < 0 || in_features <= 0 || out_features <= 0) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_backward_bias: bad sizes");
  if (rows > 0 && (!grad_y || (grad_x && !weight) || (grad_weight && !x))) return fail(MSDA_ERR_INVALID_ARG, "tc_linear_backward_bias: NULL tensor");
  bool bias_done = false;
  if (int rc = linear_backward_dispatch(static_cast<cudaStream_t>(stream), grad_y, x, weight, rows, in_features, out_features, grad_x, grad_weight,
                                        grad_bias, &bias_done)) return rc;
  if (grad_bias && !bias_done) return tc_linear_bias_grad(stream, grad_y, rows, out_features, grad_bias);
  return 0;
}

// ------------------------------------------------------------------------------- host-buffer entries
// Three streams form a pipeline per call: H2D copies -> kernels -> D2H copies, chained with events.  In the default
// synchronous mode a call returns when its results are in host memory.  With option "host_async" = 1 calls only
// enqueue (device buffers are bump-allocated from the arena and stay live until msda_host_sync()), so the upload of
// call i+1, the kernels of call i and the download of call i-1 overlap and PCIe runs full duplex.
// The arena is a ring: msda_host_fence() marks "everything enqueued so far", msda_host_wait(ticket) blocks until that
// work's results are in host memory and frees its part of the ring -- so a caller can keep two steps in flight (the
// uploads of step i+1's forward run under the downloads of step i's backward) without ever draining the pipeline.
namespace {
struct HostFence { cudaEvent_t ev; size_t head; int64_t id; };
struct HostPipe {
  std::mutex mu;
  int device = -1;
  char* base = nullptr;
  size_t cap = 0, head = 0, tail = 0;          // live bytes: [tail, head), wrapped when head < tail
  std::deque<HostFence> fences;                // oldest first
  std::vector<cudaEvent_t> fence_pool;
  int64_t fence_next = 1, fence_done = 0;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  static constexpr int kEvents = 64;
  cudaEvent_t ev_in[kEvents] = {}, ev_run[kEvents] = {};
  int ev_next = 0;
  int first_error = 0;
};
HostPipe g_pipe;

// Carves 256-byte aligned sub-buffers out of the arena, starting at the current bump offset.
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t at = off; off += (bytes + 255) & ~size_t(255); return at; }
};

int pipe_drain() {
  int rc = 0;
  if (g_pipe.s_in) rc |= check_cuda(cudaStreamSynchronize(g_pipe.s_in), "cudaStreamSynchronize(h2d)");
  if (g_pipe.s_run) rc |= check_cuda(cudaStreamSynchronize(g_pipe.s_run), "cudaStreamSynchronize(compute)");
  if (g_pipe.s_out) rc |= check_cuda(cudaStreamSynchronize(g_pipe.s_out), "cudaStreamSynchronize(d2h)");
  g_pipe.head = g_pipe.tail = 0;
  for (const HostFence& f : g_pipe.fences) { g_pipe.fence_pool.push_back(f.ev); g_pipe.fence_done = f.id; }
  g_pipe.fences.clear();
  return rc ? MSDA_ERR_CUDA : 0;
}

// the oldest fence has completed (or is waited for): its part of the ring is free again
int pipe_retire_oldest(bool block) {
  HostFence& f = g_pipe.fences.front();
  if (block) {
    if (int rc = check_cuda(cudaEventSynchronize(f.ev), "cudaEventSynchronize(fence)")) return rc;
  } else if (cudaEventQuery(f.ev) != cudaSuccess) {
    cudaGetLastError();
    return 1;                                            // still running
  }
  g_pipe.tail = f.head;
  g_pipe.fence_done = f.id;
  g_pipe.fence_pool.push_back(f.ev);
  g_pipe.fences.pop_front();
  if (g_pipe.fences.empty() && g_pipe.tail == g_pipe.head) g_pipe.head = g_pipe.tail = 0;
  return 0;
}

bool pipe_try_alloc(size_t bytes, size_t* at) {
  if (g_pipe.head >= g_pipe.tail) {
    if (g_pipe.head + bytes <= g_pipe.cap) { *at = g_pipe.head; g_pipe.head += bytes; return true; }
    if (bytes < g_pipe.tail) { *at = 0; g_pipe.head = bytes; return true; }       // wrap; the end of the ring stays unused this lap
    return false;
  }
  if (g_pipe.head + bytes < g_pipe.tail) { *at = g_pipe.head; g_pipe.head += bytes; return true; }
  return false;
}

void pipe_free() {
  if (g_pipe.base) { cudaFree(g_pipe.base); g_pipe.base = nullptr; }
  g_pipe.cap = g_pipe.head = g_pipe.tail = 0;
  for (const HostFence& f : g_pipe.fences) cudaEventDestroy(f.ev);
  g_pipe.fences.clear();
  for (cudaEvent_t e : g_pipe.fence_pool) cudaEventDestroy(e);
  g_pipe.fence_pool.clear();
  for (cudaStream_t* s : {&g_pipe.s_in, &g_pipe.s_run, &g_pipe.s_out})
    if (*s) { cudaStreamDestroy(*s); *s = nullptr; }
  for (int i = 0; i < HostPipe::kEvents; ++i) {
    if (g_pipe.ev_in[i]) { cudaEventDestroy(g_pipe.ev_in[i]); g_pipe.ev_in[i] = nullptr; }
    if (g_pipe.ev_run[i]) { cudaEventDestroy(g_pipe.ev_run[i]); g_pipe.ev_run[i] = nullptr; }
  }
}

// Reserve `bytes` of device memory for one call; returns the arena offset of its first byte in *at.
int pipe_reserve(int device, size_t bytes, size_t* at) {
  if (int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return rc;
  if (g_pipe.device != device) { pipe_drain(); pipe_free(); g_pipe.device = device; }
  if (!g_pipe.s_in) {
    for (cudaStream_t* s : {&g_pipe.s_in, &g_pipe.s_run, &g_pipe.s_out})
      if (int rc = check_cuda(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking), "cudaStreamCreate")) return rc;
    for (int i = 0; i < HostPipe::kEvents; ++i) {
      if (int rc = check_cuda(cudaEventCreateWithFlags(&g_pipe.ev_in[i], cudaEventDisableTiming), "cudaEventCreate")) return rc;
      if (int rc = check_cuda(cudaEventCreateWithFlags(&g_pipe.ev_run[i], cudaEventDisableTiming), "cudaEventCreate")) return rc;
    }
  }
  const bool async = g_opt.host_async.load() != 0;
  if (!async && g_pipe.fences.empty()) g_pipe.head = g_pipe.tail = 0;      // synchronous mode: the previous call has completed
  bytes = (bytes + 255) & ~size_t(255);
  if (pipe_try_alloc(bytes, at)) return 0;
  while (!g_pipe.fences.empty()) {                        // ring full: completed fences first, then wait for the oldest ones
    if (int rc = pipe_retire_oldest(true)) return rc;
    if (pipe_try_alloc(bytes, at)) return 0;
  }
  if (int rc = pipe_drain()) return rc;                   // nothing in flight any more, ring empty
  if (bytes > g_pipe.cap) {
    if (g_pipe.base) { cudaFree(g_pipe.base); g_pipe.base = nullptr; g_pipe.cap = 0; }
    // async: room for many calls in flight (a training clip's 73 calls stage ~7 GB; the part has 180 GB)
    const size_t want = async ? (bytes * 8 > (size_t(8) << 30) ? bytes * 8 : (size_t(8) << 30)) : bytes + bytes / 4;
    if (cudaMalloc(&g_pipe.base, want) == cudaSuccess) g_pipe.cap = want;
    else {
      cudaGetLastError();
      if (int rc = check_cuda(cudaMalloc(&g_pipe.base, bytes), "cudaMalloc(host arena)")) return rc;
      g_pipe.cap = bytes;
    }
  }
  if (!pipe_try_alloc(bytes, at)) return fail(MSDA_ERR_CUDA, "host arena: cannot place %zu bytes", bytes);
  return 0;
}

// Event pair of the next call slot (the ring is long enough: a slot is reused 64 calls later, and every call waits
// for its own events on the consuming stream before they can be re-recorded meaningfully).
void pipe_events(cudaEvent_t* in, cudaEvent_t* run) {
  const int i = g_pipe.ev_next++ % HostPipe::kEvents;
  *in = g_pipe.ev_in[i];
  *run = g_pipe.ev_run[i];
}

int pipe_finish_call() {
  if (g_opt.host_async.load() != 0) return 0;
  return check_cuda(cudaStreamSynchronize(g_pipe.s_out), "cudaStreamSynchronize(d2h)");
}

#define H2D(dst, src, bytes, what) \
  if ((bytes) > 0) if (int rc_ = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_pipe.s_in), "H2D " what)) return rc_
#define D2H(dst, src, bytes, what) \
  if ((bytes) > 0) if (int rc_ = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_pipe.s_out), "D2H " what)) return rc_
}  // namespace

int msda_forward_host(int device, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                      const void* loc, const void* aw, int N, int S, int M, int D, int L, int Lq, int P, void* out) {
  if (int rc = validate("msda_forward_host", dtype, value, shapes, level_start, loc, aw, N, S, M, D, L, Lq, P)) return rc;
  const size_t es = dtype_size(dtype), ls = loc_dtype_size(dtype);
  const size_t b_val = (size_t)N * S * M * D * es, b_loc = (size_t)N * Lq * M * L * P * 2 * ls, b_aw = b_loc / 2;
  const size_t b_out = (size_t)N * Lq * M * D * es, b_shp = (size_t)L * 2 * 8, b_lsi = (size_t)L * 8;
  if (b_out == 0) return 0;
  if (!out) return fail(MSDA_ERR_INVALID_ARG, "msda_forward_host: out is NULL");
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  Carver cv;
  const size_t o_val = cv.take(b_val), o_loc = cv.take(b_loc), o_aw = cv.take(b_aw), o_out = cv.take(b_out);
  const size_t o_shp = cv.take(b_shp), o_lsi = cv.take(b_lsi);
  size_t at = 0;
  if (int rc = pipe_reserve(device, cv.off, &at)) return rc;
  char* d = g_pipe.base + at;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(d + o_val, value, b_val, "value");
  H2D(d + o_loc, loc, b_loc, "loc");
  H2D(d + o_aw, aw, b_aw, "aw");
  H2D(d + o_shp, shapes, b_shp, "shapes");
  H2D(d + o_lsi, level_start, b_lsi, "level_start");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  if (int rc = msda_forward(g_pipe.s_run, dtype, d + o_val, reinterpret_cast<int64_t*>(d + o_shp), reinterpret_cast<int64_t*>(d + o_lsi),
                            d + o_loc, d + o_aw, N, S, M, D, L, Lq, P, d + o_out)) return rc;
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(out, d + o_out, b_out, "out");
  return pipe_finish_call();
}

int msda_backward_host(int device, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                       const void* loc, const void* aw, const void* grad_out, int N, int S, int M, int D, int L, int Lq,
                       int P, void* grad_value, void* grad_loc, void* grad_aw) {
  if (int rc = validate("msda_backward_host", dtype, value, shapes, level_start, loc, aw, N, S, M, D, L, Lq, P)) return rc;
  const size_t es = dtype_size(dtype), ls = loc_dtype_size(dtype);
  const size_t b_val = (size_t)N * S * M * D * es, b_loc = (size_t)N * Lq * M * L * P * 2 * ls, b_aw = b_loc / 2;
  const size_t b_go = (size_t)N * Lq * M * D * es, b_shp = (size_t)L * 2 * 8, b_lsi = (size_t)L * 8;
  const size_t b_ws = msda_backward_workspace_bytes(dtype, N, S, M, D);
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  Carver cv;
  const size_t o_val = cv.take(b_val), o_loc = cv.take(b_loc), o_aw = cv.take(b_aw), o_go = cv.take(b_go);
  const size_t o_shp = cv.take(b_shp), o_lsi = cv.take(b_lsi);
  const size_t o_gv = cv.take(b_val), o_gl = cv.take(b_loc), o_ga = cv.take(b_aw), o_ws = cv.take(b_ws);
  size_t at = 0;
  if (int rc = pipe_reserve(device, cv.off, &at)) return rc;
  char* d = g_pipe.base + at;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(d + o_val, value, b_val, "value");
  H2D(d + o_loc, loc, b_loc, "loc");
  H2D(d + o_aw, aw, b_aw, "aw");
  H2D(d + o_go, grad_out, b_go, "grad_out");
  H2D(d + o_shp, shapes, b_shp, "shapes");
  H2D(d + o_lsi, level_start, b_lsi, "level_start");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  if (int rc = msda_backward(g_pipe.s_run, dtype, d + o_val, reinterpret_cast<int64_t*>(d + o_shp), reinterpret_cast<int64_t*>(d + o_lsi),
                             d + o_loc, d + o_aw, d + o_go, N, S, M, D, L, Lq, P, d + o_gv, d + o_gl, d + o_ga,
                             b_ws ? d + o_ws : nullptr, b_ws)) return rc;
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(grad_value, d + o_gv, b_val, "grad_value");
  D2H(grad_loc, d + o_gl, b_loc, "grad_loc");
  D2H(grad_aw, d + o_ga, b_aw, "grad_aw");
  return pipe_finish_call();
}

int mask_logits_forward_host(int device, int in_dtype, int out_dtype, const void* coeff, const void* proto, int B, int Q,
                             int K, int64_t Ncols, void* out) {
  if (B < 0 || Q < 0 || K <= 0 || Ncols < 0) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward_host: bad sizes");
  const size_t ei = dtype_size(in_dtype), eo = dtype_size(out_dtype);
  const size_t b_c = (size_t)B * Q * K * ei, b_p = (size_t)B * K * Ncols * ei, b_o = (size_t)B * Q * Ncols * eo;
  if (b_o == 0) return 0;
  if (!coeff || !proto || !out) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward_host: NULL tensor");
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  Carver cv;
  const size_t o_c = cv.take(b_c), o_p = cv.take(b_p), o_o = cv.take(b_o);
  size_t at = 0;
  if (int rc = pipe_reserve(device, cv.off, &at)) return rc;
  char* d = g_pipe.base + at;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(d + o_c, coeff, b_c, "coeff");
  H2D(d + o_p, proto, b_p, "proto");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  if (int rc = mask_logits_forward(g_pipe.s_run, in_dtype, out_dtype, d + o_c, d + o_p, B, Q, K, Ncols, d + o_o)) return rc;
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(out, d + o_o, b_o, "out");
  return pipe_finish_call();
}

// ---- "saved" host entries: autograd's save_for_backward for the host-buffer ABI.  The forward keeps the device copies of
// its inputs in a pooled block and hands out a handle; the backward uploads only grad_out.  Blocks are recycled through a
// grow-only pool (no cudaMalloc / cudaFree in steady state); reuse is ordered after the last kernel that read the block.
namespace {
struct SavedBlock {
  char* base = nullptr;
  size_t cap = 0;
  cudaEvent_t last_use = nullptr;
  bool in_use = false;
  uint32_t gen = 1;
  int kind = 0;                              // 1 = msda (plain / grouped), 2 = mask contraction
  int dtype = 0, out_dtype = 0, N = 0, S = 0, M = 0, D = 0, G = 1, L = 0, Lq = 0, P = 0, B = 0, Q = 0, K = 0;
  int64_t Ncols = 0;
  float scale = 1.f;
  size_t o_a = 0, o_b = 0, o_c = 0, o_shp = 0, o_lsi = 0;   // msda: value / loc / aw;  mask: coeff / proto
};
std::vector<SavedBlock*> g_saved;
std::deque<int> g_saved_free;                 // free blocks with memory, oldest release first
size_t g_saved_bytes = 0;
constexpr size_t kSavedPoolLimit = size_t(24) << 30;

// Blocks are released in the order their last readers were enqueued on the compute stream, so the oldest free block that
// fits is the one whose readers finish first.  If even that one is still being read (two steps in flight: the next step's
// forward is enqueued while this step's backward runs) the pool grows instead of making the upload wait, up to a limit.
int saved_acquire(size_t bytes, int* index) {
  const size_t mb = size_t(1) << 20;
  const size_t want = (bytes + mb - 1) & ~(mb - 1);
  const size_t slack = want > 4 * mb ? want : 4 * mb;
  int pick = -1;
  bool pick_done = false;
  std::deque<int>::iterator pick_it;
  for (auto it = g_saved_free.begin(); it != g_saved_free.end(); ++it) {
    SavedBlock* b = g_saved[*it];
    if (b->cap < bytes || b->cap > want + slack) continue;
    pick = *it;
    pick_it = it;
    pick_done = cudaEventQuery(b->last_use) == cudaSuccess;
    if (!pick_done) cudaGetLastError();
    break;
  }
  if (pick >= 0 && (pick_done || g_saved_bytes + want > kSavedPoolLimit)) {
    g_saved_free.erase(pick_it);
  } else {
    while (g_saved_bytes + want > kSavedPoolLimit && !g_saved_free.empty()) {      // make room: drop the oldest free blocks
      SavedBlock* b = g_saved[g_saved_free.front()];
      g_saved_free.pop_front();
      cudaEventSynchronize(b->last_use);
      cudaFree(b->base);
      g_saved_bytes -= b->cap;
      b->base = nullptr;
      b->cap = 0;
    }
    pick = -1;
    for (size_t i = 0; i < g_saved.size(); ++i)
      if (!g_saved[i]->in_use && !g_saved[i]->base) { pick = static_cast<int>(i); break; }
    if (pick < 0) { g_saved.push_back(new SavedBlock()); pick = static_cast<int>(g_saved.size()) - 1; }
    SavedBlock* b = g_saved[pick];
    if (int rc = check_cuda(cudaMalloc(&b->base, want), "cudaMalloc(saved block)")) { b->base = nullptr; return rc; }
    b->cap = want;
    g_saved_bytes += want;
    if (!b->last_use)
      if (int rc = check_cuda(cudaEventCreateWithFlags(&b->last_use, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    cudaEventRecord(b->last_use, g_pipe.s_run);
  }
  SavedBlock* b = g_saved[pick];
  b->in_use = true;
  cudaStreamWaitEvent(g_pipe.s_in, b->last_use, 0);            // uploads into the block wait for its previous readers
  *index = pick;
  return 0;
}

SavedBlock* saved_lookup(const char* who, int64_t handle, int kind) {
  const uint32_t index = static_cast<uint32_t>(handle & 0xffffffff), gen = static_cast<uint32_t>(static_cast<uint64_t>(handle) >> 32);
  if (handle <= 0 || index >= g_saved.size() || !g_saved[index]->in_use || g_saved[index]->gen != gen || g_saved[index]->kind != kind) {
    fail(MSDA_ERR_INVALID_ARG, "%s: stale or foreign handle (each handle is consumed by one backward / release)", who);
    return nullptr;
  }
  return g_saved[index];
}

void saved_free(SavedBlock* b) {                // call after recording b->last_use behind the block's last reader
  b->in_use = false;
  ++b->gen;
  for (size_t i = 0; i < g_saved.size(); ++i)
    if (g_saved[i] == b) { g_saved_free.push_back(static_cast<int>(i)); break; }
}

// make sure the pipeline (streams, events) exists on `device` without reserving arena space
int pipe_open(int device) {
  size_t at = 0;
  return pipe_reserve(device, 0, &at);
}
}  // namespace

int msda_forward_host_saved(int device, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                            const void* loc, const void* aw, int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                            void* out, int64_t* saved) {
  if (int rc = validate("msda_forward_host_saved", dtype, value, shapes, level_start, loc, aw, N, S, M, D, L, Lq, P)) return rc;
  if (!saved) return fail(MSDA_ERR_INVALID_ARG, "msda_forward_host_saved: saved is NULL");
  if (G < 1) return fail(MSDA_ERR_INVALID_ARG, "msda_forward_host_saved: G=%d", G);
  *saved = 0;
  const size_t es = dtype_size(dtype), ls = loc_dtype_size(dtype);
  const size_t b_val = (size_t)N * S * M * D * es, b_loc = (size_t)N * Lq * M * L * P * 2 * ls, b_aw = b_loc / 2;
  const size_t b_out = (size_t)N * Lq * M * D * es, b_shp = (size_t)G * L * 2 * 8, b_lsi = (size_t)G * L * 8;
  if (b_out > 0 && !out) return fail(MSDA_ERR_INVALID_ARG, "msda_forward_host_saved: out is NULL");
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  size_t at = 0;
  if (int rc = pipe_reserve(device, b_out, &at)) return rc;
  char* d_out = g_pipe.base + at;
  Carver cv;
  const size_t o_val = cv.take(b_val), o_loc = cv.take(b_loc), o_aw = cv.take(b_aw), o_shp = cv.take(b_shp), o_lsi = cv.take(b_lsi);
  int index = -1;
  if (int rc = saved_acquire(cv.off + 256, &index)) return rc;
  SavedBlock* b = g_saved[index];
  b->kind = 1; b->dtype = dtype; b->N = N; b->S = S; b->M = M; b->D = D; b->G = G; b->L = L; b->Lq = Lq; b->P = P; b->scale = scale;
  b->o_a = o_val; b->o_b = o_loc; b->o_c = o_aw; b->o_shp = o_shp; b->o_lsi = o_lsi;
  char* d = b->base;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(d + o_val, value, b_val, "value");
  H2D(d + o_loc, loc, b_loc, "loc");
  H2D(d + o_aw, aw, b_aw, "aw");
  H2D(d + o_shp, shapes, b_shp, "shapes");
  H2D(d + o_lsi, level_start, b_lsi, "level_start");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  int rc = 0;
  if (b_out > 0)
    rc = forward_impl("msda_forward_host_saved", g_pipe.s_run, dtype, d + o_val, reinterpret_cast<int64_t*>(d + o_shp),
                      reinterpret_cast<int64_t*>(d + o_lsi), d + o_loc, d + o_aw, N, S, M, D, G, L, Lq, P, scale, d_out);
  cudaEventRecord(b->last_use, g_pipe.s_run);
  if (rc) { saved_free(b); return rc; }
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(out, d_out, b_out, "out");
  *saved = (static_cast<int64_t>(b->gen) << 32) | static_cast<int64_t>(index);
  return pipe_finish_call();
}

int msda_backward_host_saved(int64_t saved, const void* grad_out, void* grad_value, void* grad_loc, void* grad_aw) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  SavedBlock* b = saved_lookup("msda_backward_host_saved", saved, 1);
  if (!b) return MSDA_ERR_INVALID_ARG;
  const size_t es = dtype_size(b->dtype), ls = loc_dtype_size(b->dtype);
  const size_t b_val = (size_t)b->N * b->S * b->M * b->D * es, b_loc = (size_t)b->N * b->Lq * b->M * b->L * b->P * 2 * ls, b_aw = b_loc / 2;
  const size_t b_go = (size_t)b->N * b->Lq * b->M * b->D * es;
  const size_t b_ws = msda_backward_workspace_bytes(b->dtype, b->N, b->S, b->M, b->D);
  if ((b_go > 0 && !grad_out) || (b_val > 0 && !grad_value) || (b_loc > 0 && (!grad_loc || !grad_aw)))
    return fail(MSDA_ERR_INVALID_ARG, "msda_backward_host_saved: NULL tensor");
  Carver cv;
  const size_t o_go = cv.take(b_go), o_gv = cv.take(b_val), o_gl = cv.take(b_loc), o_ga = cv.take(b_aw), o_ws = cv.take(b_ws);
  size_t at = 0;
  if (int rc = pipe_reserve(g_pipe.device, cv.off, &at)) return rc;
  char* t = g_pipe.base + at;
  char* d = b->base;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(t + o_go, grad_out, b_go, "grad_out");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  const int rc = backward_impl("msda_backward_host_saved", g_pipe.s_run, b->dtype, d + b->o_a, reinterpret_cast<int64_t*>(d + b->o_shp),
                               reinterpret_cast<int64_t*>(d + b->o_lsi), d + b->o_b, d + b->o_c, t + o_go, b->N, b->S, b->M, b->D, b->G,
                               b->L, b->Lq, b->P, b->scale, t + o_gv, t + o_gl, t + o_ga, b_ws ? t + o_ws : nullptr, b_ws);
  cudaEventRecord(b->last_use, g_pipe.s_run);
  saved_free(b);
  if (rc) return rc;
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(grad_value, t + o_gv, b_val, "grad_value");
  D2H(grad_loc, t + o_gl, b_loc, "grad_loc");
  D2H(grad_aw, t + o_ga, b_aw, "grad_aw");
  return pipe_finish_call();
}

int mask_logits_forward_host_saved(int device, int in_dtype, int out_dtype, const void* coeff, const void* proto, int B, int Q,
                                   int K, int64_t Ncols, void* out, int64_t* saved) {
  if (B < 0 || Q < 0 || K <= 0 || Ncols < 0) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward_host_saved: bad sizes");
  if (!saved) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward_host_saved: saved is NULL");
  *saved = 0;
  const size_t ei = dtype_size(in_dtype), eo = dtype_size(out_dtype);
  const size_t b_c = (size_t)B * Q * K * ei, b_p = (size_t)B * K * Ncols * ei, b_o = (size_t)B * Q * Ncols * eo;
  if ((b_c > 0 && !coeff) || (b_p > 0 && !proto) || (b_o > 0 && !out)) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_forward_host_saved: NULL tensor");
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  size_t at = 0;
  if (int rc = pipe_reserve(device, b_o, &at)) return rc;
  char* d_out = g_pipe.base + at;
  Carver cv;
  const size_t o_c = cv.take(b_c), o_p = cv.take(b_p);
  int index = -1;
  if (int rc = saved_acquire(cv.off + 256, &index)) return rc;
  SavedBlock* b = g_saved[index];
  b->kind = 2; b->dtype = in_dtype; b->out_dtype = out_dtype; b->B = B; b->Q = Q; b->K = K; b->Ncols = Ncols; b->o_a = o_c; b->o_b = o_p;
  char* d = b->base;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(d + o_c, coeff, b_c, "coeff");
  H2D(d + o_p, proto, b_p, "proto");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  int rc = 0;
  if (b_o > 0) rc = mask_logits_forward(g_pipe.s_run, in_dtype, out_dtype, d + o_c, d + o_p, B, Q, K, Ncols, d_out);
  cudaEventRecord(b->last_use, g_pipe.s_run);
  if (rc) { saved_free(b); return rc; }
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  D2H(out, d_out, b_o, "out");
  *saved = (static_cast<int64_t>(b->gen) << 32) | static_cast<int64_t>(index);
  return pipe_finish_call();
}

int mask_logits_backward_host_saved(int64_t saved, const void* grad_out, void* grad_coeff, void* grad_proto) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  SavedBlock* b = saved_lookup("mask_logits_backward_host_saved", saved, 2);
  if (!b) return MSDA_ERR_INVALID_ARG;
  const size_t e = dtype_size(b->dtype);
  const size_t b_c = (size_t)b->B * b->Q * b->K * e, b_p = (size_t)b->B * b->K * b->Ncols * e, b_go = (size_t)b->B * b->Q * b->Ncols * e;
  if (b_go > 0 && !grad_out) return fail(MSDA_ERR_INVALID_ARG, "mask_logits_backward_host_saved: grad_out is NULL");
  Carver cv;
  const size_t o_go = cv.take(b_go), o_gc = cv.take(b_c), o_gp = cv.take(b_p);
  size_t at = 0;
  if (int rc = pipe_reserve(g_pipe.device, cv.off, &at)) return rc;
  char* t = g_pipe.base + at;
  char* d = b->base;
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  H2D(t + o_go, grad_out, b_go, "grad_out");
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaStreamWaitEvent(g_pipe.s_run, e_in, 0);
  const int rc = mask_logits_backward(g_pipe.s_run, b->dtype, d + b->o_a, d + b->o_b, t + o_go, b->B, b->Q, b->K, b->Ncols,
                                      grad_coeff ? t + o_gc : nullptr, grad_proto ? t + o_gp : nullptr);
  cudaEventRecord(b->last_use, g_pipe.s_run);
  saved_free(b);
  if (rc) return rc;
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  if (grad_coeff) D2H(grad_coeff, t + o_gc, b_c, "grad_coeff");
  if (grad_proto) D2H(grad_proto, t + o_gp, b_p, "grad_proto");
  return pipe_finish_call();
}

int msda_host_saved_release(int64_t saved) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  const uint32_t index = static_cast<uint32_t>(saved & 0xffffffff), gen = static_cast<uint32_t>(static_cast<uint64_t>(saved) >> 32);
  if (saved <= 0 || index >= g_saved.size() || !g_saved[index]->in_use || g_saved[index]->gen != gen)
    return fail(MSDA_ERR_INVALID_ARG, "msda_host_saved_release: stale handle");
  saved_free(g_saved[index]);
  return 0;
}

int msda_host_sync(void) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (g_pipe.device < 0) return 0;
  cudaSetDevice(g_pipe.device);
  return pipe_drain();
}

int msda_host_fence(int64_t* ticket) {
  if (!ticket) return fail(MSDA_ERR_INVALID_ARG, "msda_host_fence: ticket is NULL");
  *ticket = 0;
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (g_pipe.device < 0 || !g_pipe.s_out) return 0;      // nothing was ever enqueued: ticket 0 is always complete
  cudaSetDevice(g_pipe.device);
  cudaEvent_t ev = nullptr;
  if (!g_pipe.fence_pool.empty()) { ev = g_pipe.fence_pool.back(); g_pipe.fence_pool.pop_back(); }
  else if (int rc = check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate(fence)")) return rc;
  // every call's downloads wait for its kernels, which wait for its uploads, and the three streams are in order:
  // an event behind the last download (plus the other two streams, for calls without outputs) covers all earlier work
  cudaEvent_t e_in, e_run;
  pipe_events(&e_in, &e_run);
  cudaEventRecord(e_in, g_pipe.s_in);
  cudaEventRecord(e_run, g_pipe.s_run);
  cudaStreamWaitEvent(g_pipe.s_out, e_in, 0);
  cudaStreamWaitEvent(g_pipe.s_out, e_run, 0);
  if (int rc = check_cuda(cudaEventRecord(ev, g_pipe.s_out), "cudaEventRecord(fence)")) { g_pipe.fence_pool.push_back(ev); return rc; }
  g_pipe.fences.push_back(HostFence{ev, g_pipe.head, g_pipe.fence_next});
  *ticket = g_pipe.fence_next++;
  return 0;
}

int msda_host_wait(int64_t ticket) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (ticket < 0 || ticket >= g_pipe.fence_next) return fail(MSDA_ERR_INVALID_ARG, "msda_host_wait: unknown ticket %lld", (long long)ticket);
  if (g_pipe.device >= 0) cudaSetDevice(g_pipe.device);
  while (!g_pipe.fences.empty() && g_pipe.fences.front().id <= ticket)
    if (int rc = pipe_retire_oldest(true)) return rc;
  return 0;                                               // tickets at or below fence_done were completed by a wait or a sync
}

int msda_host_arena_release(void) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (g_pipe.device >= 0) {
    cudaSetDevice(g_pipe.device);
    pipe_drain();
    for (SavedBlock* b : g_saved) {                       // outstanding handles become stale
      if (b->base) cudaFree(b->base);
      if (b->last_use) cudaEventDestroy(b->last_use);
      delete b;
    }
    g_saved.clear();
    g_saved_free.clear();
    g_saved_bytes = 0;
    pipe_free();
  }
  g_pipe.device = -1;
  return 0;
}

}  // extern "C"
