// Internal glue shared by the translation units of libmsda_b200.so (not part of the public ABI).
#pragma once

#include <atomic>

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {

// Records `msg` as the calling thread's last error and returns `code`.
int fail(int code, const char* fmt, ...);
// Converts a CUDA error into MSDA_ERR_CUDA (+ message); returns 0 when err == cudaSuccess.
int check_cuda(cudaError_t err, const char* what);
// Call after every kernel launch: counts it and turns launch errors into status codes.
int after_launch(const char* kernel_name);

int option(const char* key);          // current value of a tuning knob

// cudaFuncSetAttribute is per device (context), so "set once per process" leaves every device but the first without its
// opt-in to > 48 KB of dynamic shared memory (launches then fail with `invalid argument` under nn.DataParallel, per-device
// threads, or the *_host entries' `device` argument).  This remembers (function, attribute, current device) triples instead.
int ensure_func_attr_impl(const void* func, cudaFuncAttribute attr, int value);
template <typename K>
inline int ensure_func_attr(K kernel, cudaFuncAttribute attr, int value) {
  return ensure_func_attr_impl(reinterpret_cast<const void*>(kernel), attr, value);
}
int sm_count();                       // multiprocessors of the CURRENT device (cached per device)

struct Options {
  std::atomic<int> fwd_variant{0};    // 0 = auto (fast2 / fast when eligible), 1 = force generic, 2 = first-generation fast, 3 = fast2 with batched gathers (80 registers)
  std::atomic<int> bwd_variant{0};    // 0 = auto (fast2 / fast), 1 = generic, 2 = first-generation fast
  std::atomic<int> chunk_pairs{0};    // 0 = auto
  std::atomic<int> mask_variant{0};   // 0 = auto (tcgen05 when eligible), 1 = SIMT fp32, 2 = force tcgen05
  std::atomic<int> mask_debug{0};     // 1 = the tcgen05 mask kernel records per-item clock stamps of CTA 0
  std::atomic<int> host_async{0};     // 1 = the *_host entries only enqueue; msda_host_sync() completes them
  std::atomic<int> profile{0};        // 1 = bracket the main kernels with CUDA events (msda_profile_read)
  std::atomic<int> consumer_tc{0};    // matcher mask costs: 0 = auto (tensor-core kernel when eligible), 1 = SIMT kernel, 2 = require the tensor-core kernel
  std::atomic<int> consumer_ctas{0};  // > 0: CTAs of the matcher-cost / IoU kernels (tuning; 0 = auto)
  std::atomic<int> pdl{1};            // 1 = launch the sampling kernels with programmatic stream serialization (see msda_launch.cuh)
  std::atomic<int> pair_map{0};       // order in which a CTA of the fast2 sampling kernels walks its pairs: 0 = auto, 1 = linear (chunk / M queries x all heads), 2 = head-run (one head x chunk queries; PairMap in msda_fast2.cuh)
  std::atomic<int> mask_a_tmem{0};    // 3xTF32 mask kernels (mask_tc4.cuh): plane operand through tensor memory: 0 = on, 1 = off (A/B)
  std::atomic<int> gemm_stream_k{1};  // Linear-layer GEMMs: 1 = (tile, chunk) units dealt out as one contiguous range per CTA where that pays (launch_gemm3x), 2 = always, 0 = whole tiles round-robin (A/B; gemm3x.cuh)
  std::atomic<int> bwd_merge{1};      // 1 = merge grad_value reductions of a (pair, level) that hit the same row (P = 2 or 4); 0 = off (A/B)
};
const Options& options();


// RAII event bracket around one kernel launch; active only when the "profile" option is 1 and the
// stream is not being captured into a CUDA graph.
class ProfScope {
 public:
  ProfScope(cudaStream_t st, int kind, int64_t units);
  ~ProfScope();
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;
 private:
  cudaStream_t st_;
  int kind_;
  int64_t units_;
  cudaEvent_t a_ = nullptr, b_ = nullptr;
  bool on_ = false;
};

// mask_gemm.cu
int mask_debug_copy(long long* host80);
int mask_forward_dispatch(cudaStream_t stream, int in_dtype, int out_dtype, const void* coeff, const void* proto,
                          int B, int Q, int K, int64_t Ncols, void* out);
int mask_backward_dispatch(cudaStream_t stream, int dtype, const void* coeff, const void* proto, const void* grad_out,
                           int B, int Q, int K, int64_t Ncols, void* grad_coeff, void* grad_proto);

size_t match_cost_tc_workspace_floats();
bool match_cost_tc_eligible(const void* coeff, const void* proto, const void* tgt, int K, int64_t Ncols);
int match_cost_tc_dispatch(cudaStream_t stream, const float* coeff, const float* proto, const float* tgt, int Q, int K, int G, int64_t Ncols,
                           float* ws, float* cost_bce, float* cost_dice, int ld);

int linear_forward_dispatch(cudaStream_t stream, const void* x, const void* w, const void* bias, const unsigned char* row_mask,
                            int64_t rows, int in_f, int out_f, void* y);
int64_t gemm_flag_timeouts();
int linear_forward_packed_dispatch(cudaStream_t stream, const void* x, const void* w, const void* bias, const unsigned char* row_mask, int N,
                                   int S, int in_f, int heads, const int64_t* shapes, const int64_t* level_start, int L, void* packed);
int linear_backward_dispatch(cudaStream_t stream, const void* gy, const void* x, const void* w, int64_t rows, int in_f, int out_f,
                             void* gx, void* gw, void* gb = nullptr, bool* gb_done = nullptr);

inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case MSDA_F32: return 4;
    case MSDA_BF16: return 2;
    case MSDA_F64: return 8;
    case MSDA_BF16_LOC32: return 2;
    case MSDA_F16: return 2;              // mask contraction only
    default: return 0;
  }
}
inline size_t loc_dtype_size(int dtype) {
  switch (dtype) {
    case MSDA_F32: return 4;
    case MSDA_BF16: return 2;
    case MSDA_F64: return 8;
    case MSDA_BF16_LOC32: return 4;
    default: return 0;
  }
}

}  // namespace msda
