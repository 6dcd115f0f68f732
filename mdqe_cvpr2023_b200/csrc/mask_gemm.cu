// Mask contraction dispatch: out[b,q,n] = sum_k coeff[b,q,k] * proto[b,k,n] and its gradients.
// See include/msda_b200.h (mask_logits_*) for the contract.
#include "mask_simt.cuh"
#include "msda_internal.h"

namespace msda {

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

int mask_forward_dispatch(cudaStream_t st, int in_dtype, int out_dtype, const void* coeff, const void* proto, int B,
                          int Q, int K, int64_t Ncols, void* out) {
  if (B > 65535) return fail(MSDA_ERR_UNSUPPORTED, "mask_logits_forward: B=%d > 65535", B);
  if (in_dtype == MSDA_F32 && out_dtype == MSDA_F32) return launch_mask_simt<float, float>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_F32 && out_dtype == MSDA_BF16) return launch_mask_simt<float, __nv_bfloat16>(st, coeff, proto, out, B, Q, K, Ncols);
  if (in_dtype == MSDA_BF16 && out_dtype == MSDA_F32) return launch_mask_simt<__nv_bfloat16, float>(st, coeff, proto, out, B, Q, K, Ncols);
  return launch_mask_simt<__nv_bfloat16, __nv_bfloat16>(st, coeff, proto, out, B, Q, K, Ncols);
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
  const dim3 grid(static_cast<unsigned>((Ncols + kMaskTN - 1) / kMaskTN), (K + kMaskKC - 1) / kMaskKC, B);
  ProfScope prof(st, MSDA_PROF_MASK_BWD, (int64_t)B * Q * Ncols);
  mask_bwd_simt_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(coeff), static_cast<const float*>(proto),
                                             static_cast<const float*>(grad_out), static_cast<float*>(grad_coeff),
                                             static_cast<float*>(grad_proto), Q, K, Ncols);
  return after_launch("mask_bwd_simt_kernel");
}

}  // namespace msda
