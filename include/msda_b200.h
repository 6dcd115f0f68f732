/*
 * msda_b200.h -- C ABI of the B200-native (sm_100a) MDQE hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / ATen types.  It is what a
 * binding for the reference's extension seam binds.  Reference interfaces replaced
 * (paths relative to /root/reference/mdqe/models/ops unless noted):
 *
 *   msda_forward        <- ms_deform_attn_forward  (src/vision.cpp:14, src/ms_deform_attn.h:20-39,
 *                          src/cuda/ms_deform_attn_cuda.cu:20-80) and the launcher
 *                          ms_deformable_im2col_cuda (src/cuda/ms_deform_im2col_cuda.cuh:923-954)
 *   msda_backward       <- ms_deform_attn_backward (src/vision.cpp:15, src/ms_deform_attn.h:41-61,
 *                          src/cuda/ms_deform_attn_cuda.cu:83-153) and ms_deformable_col2im_cuda
 *                          (src/cuda/ms_deform_im2col_cuda.cuh:956-1326)
 *   mask_logits_*       <- torch.einsum('bqm,bmthw->bqthw') at mdqe/models/transformer_dec.py:255,
 *                          mdqe/mdqe.py:384, mdqe/models/matcher.py:182, mdqe/models/criterion.py:440
 *   *_host variants     <- the same calls with HOST buffers (the library stages them through
 *                          device memory); this is the end-to-end entry bench.py times as `e2e`.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative msda_status; msda_last_error() returns a
 *     thread-local, human readable description of the last failure (the reference only printf'd
 *     launch errors: ms_deform_im2col_cuda.cuh:948-952).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *     enqueue work: no host synchronisation, no allocation, safe under CUDA-graph capture.
 *     The *_host variants synchronise the stream before returning.
 *   - All device tensors are contiguous, row-major, in the reference's layouts:
 *       value [N,S,M,D]   shapes [L,2] (H,W) int64 ON DEVICE   level_start [L] int64 ON DEVICE
 *       loc [N,Lq,M,L,P,2] (x,y) normalised to [0,1]   aw [N,Lq,M,L,P]   out / grad_out [N,Lq,M*D]
 *   - The caller owns every buffer.  There is no CPU fallback: without a CUDA device every compute
 *     entry fails with MSDA_ERR_CUDA.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 6

typedef enum {
  MSDA_OK = 0,
  MSDA_ERR_INVALID_ARG = -1,   /* NULL pointer, non-positive size, unsupported dtype ...        */
  MSDA_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels implement (e.g. L > 32)        */
  MSDA_ERR_CUDA = -3,          /* a CUDA runtime call or kernel launch failed                    */
  MSDA_ERR_WORKSPACE = -4      /* workspace missing or too small                                */
} msda_status;

/* Storage type of the tensors.  Arithmetic is fp32 (fp64 for MSDA_F64) in every mode.
 *   MSDA_F32        : everything fp32 (the reference's only mode besides fp64:
 *                     AT_DISPATCH_FLOATING_TYPES, ms_deform_attn_cuda.cu:64)
 *   MSDA_BF16       : value/out/grad_out/grad_value AND loc/aw/grad_loc/grad_aw are bf16
 *   MSDA_F64        : everything fp64 (gradcheck path of ops/test.py:63-78)
 *   MSDA_BF16_LOC32 : value/out/grad_out/grad_value bf16; loc/aw/grad_loc/grad_aw fp32
 *   MSDA_F16        : IEEE half -- mask_logits_forward only: the reference evaluates under autocast (train_net.py:207-208), so
 *                     the einsum at mdqe/mdqe.py:384 and transformer_dec.py:255 runs in fp16 there; the sampling operator
 *                     itself is cast to fp32 by its custom_fwd (ms_deform_attn_func.py:24) and has no fp16 mode
 */
typedef enum { MSDA_F32 = 0, MSDA_BF16 = 1, MSDA_F64 = 2, MSDA_BF16_LOC32 = 3, MSDA_F16 = 4 } msda_dtype;

int msda_abi_version(void);
const char* msda_last_error(void);

/* Tuning knobs (kernel variant selection used by bench.py / tests); unknown keys fail.
 * Keys: "fwd_variant", "bwd_variant", "chunk_pairs", "mask_variant", "profile", "mask_debug", "host_async",
 * "consumer_ctas", "bwd_merge" (1 = the backward merges the grad_value reductions of one query's
 * points that hit the same value row -- default; 0 = one reduction per corner, for A/B).  mask_variant: 0 auto (tensor cores when eligible), 1 SIMT, 2 require tensor cores,
 * (earlier A/B kernels were removed in round 2). */
int msda_set_option(const char* key, int value);
int msda_get_option(const char* key, int* value);

/* out[n,q,m*D+c] = sum_{l,p} aw[n,q,m,l,p] * bilinear(value_l[n,:,m,c], loc[n,q,m,l,p]*(W_l,H_l)-0.5)
 * with zero padding; `out` is fully overwritten (no pre-zeroing needed). */
int msda_forward(void* stream, int dtype,
                 const void* value, const int64_t* shapes, const int64_t* level_start,
                 const void* loc, const void* aw,
                 int N, int S, int M, int D, int L, int Lq, int P,
                 void* out);

/* Bytes of scratch msda_backward needs for this problem (0 for MSDA_F32 / MSDA_F64; the bf16
 * modes accumulate grad_value in an fp32 image of it). */
size_t msda_backward_workspace_bytes(int dtype, int N, int S, int M, int D);

/* grad_value, grad_loc, grad_aw are fully written (grad_value is zero-filled inside the call).
 * `workspace` may be NULL when msda_backward_workspace_bytes() is 0. */
int msda_backward(void* stream, int dtype,
                  const void* value, const int64_t* shapes, const int64_t* level_start,
                  const void* loc, const void* aw, const void* grad_out,
                  int N, int S, int M, int D, int L, int Lq, int P,
                  void* grad_value, void* grad_loc, void* grad_aw,
                  void* workspace, size_t workspace_bytes);

/* Grouped ("temporal") form used by the clip-level decoder attention (ms_deform_attn.py:219-235): G level tables of L
 * levels each share ONE set of sampling locations / attention weights,
 *     out = scale * sum_{g<G} msda(value, shapes[g], level_start[g], loc, aw)
 * with shapes [G,L,2] and level_start [G,L] (int64, device).  In MDQE g runs over the pyramid levels, the L "levels" are
 * the T frames of the clip inside a [B, T*S, M, D] value tensor (level_start[g][t] = t*S + start_g) and scale = 1/G.
 * One launch instead of G, no per-level copies, one zero-fill of grad_value.  Supported where the fast kernels are
 * (D in {32,24}, L*P in {8,12,16}, fp32/bf16, G*L <= 32); otherwise MSDA_ERR_UNSUPPORTED (call msda_forward per group). */
int msda_forward_grouped(void* stream, int dtype,
                         const void* value, const int64_t* shapes, const int64_t* level_start,
                         const void* loc, const void* aw,
                         int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                         void* out);
int msda_backward_grouped(void* stream, int dtype,
                          const void* value, const int64_t* shapes, const int64_t* level_start,
                          const void* loc, const void* aw, const void* grad_out,
                          int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                          void* grad_value, void* grad_loc, void* grad_aw,
                          void* workspace, size_t workspace_bytes);

/* The backward accumulates grad_value with vector reductions, so its accumulator -- grad_value itself (fp32 / fp64) or the
 * fp32 workspace (bf16 modes) -- must start at zero.  The reference zero-fills right before its kernel (`at::zeros_like`,
 * ms_deform_attn_cuda.cu:121), which puts a 20.9 MB memset per call on the critical path (6.7 % of the R50_ovis_360 training
 * step).  A caller that can fill the buffer earlier and elsewhere -- on a side stream while the forward pass runs, as
 * MSDeformAttnFunction and bench.py do -- passes MSDA_BWD_ACC_ZEROED and the call launches the sampling kernel only.
 * msda_zero_fill is that fill as a stream-ordered call (no allocation, graph-capturable).  G = 1, scale = 1: the plain operator. */
#define MSDA_BWD_ACC_ZEROED 1
int msda_zero_fill(void* stream, void* ptr, size_t bytes);
int msda_backward_grouped_flags(void* stream, int dtype,
                                const void* value, const int64_t* shapes, const int64_t* level_start,
                                const void* loc, const void* aw, const void* grad_out,
                                int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                                void* grad_value, void* grad_loc, void* grad_aw,
                                void* workspace, size_t workspace_bytes, int flags);

/* Gradient all-reduce of clip-sharded data-parallel training over NVLink peer memory (the reference leaves it to
 * PyTorch DDP over NCCL, train_net.py:256-271 -- detectron2's launch wraps the model in DistributedDataParallel).  One small
 * kernel (n_ctas CTAs) instead of NCCL's 24-32 channel CTAs, so that the bucket can be reduced UNDER the encoder backward
 * without taking its SMs:   bucket[i] = scale * sum_r bucket_r[i]   in place, in every rank's replica.
 *   peer_ptrs[world]  : the bucket's base address as mapped in THIS process for every rank (peer_ptrs[rank] = the local one);
 *                       a symmetric allocation -- same size and layout on every rank (cudaIpc / VMM handles / torch symmetric memory)
 *   multicast_ptr     : NVSwitch multicast mapping of the same allocation (algo 1) or 0
 *   flag_ptrs[world]  : msda_allreduce_flag_bytes(n_ctas) bytes per rank, mapped like the bucket, zero before the first call
 *   error_word        : int on the local device, set to 1 if a peer never arrived (bounded spin instead of a hang)
 *   algo 0            : two-shot P2P -- rank r loads slice r from every peer, adds in rank order, stores the sum to every peer
 *   algo 1            : multimem -- multimem.ld_reduce / multimem.st: the switch adds and broadcasts (needs multicast_ptr)
 * offset_elems / n_elems address a sub-range of the bucket (multiples of 4 floats).  Every rank must issue the same calls in
 * the same order; calls only enqueue (graph-capturable).  world <= msda_allreduce_max_ranks(). */
int msda_allreduce_max_ranks(void);
size_t msda_allreduce_flag_bytes(int n_ctas);
int msda_allreduce_f32(void* stream, int algo, int rank, int world, const uint64_t* peer_ptrs, uint64_t multicast_ptr,
                       const uint64_t* flag_ptrs, void* error_word, int64_t offset_elems, int64_t n_elems, float scale, int n_ctas);

/* bf16 forward on a paired-corner value layout (north_star: bf16x2 loads of a value tensor re-laid per level).  The
 * forward is bound by the number of 128-byte lines it gathers -- four per sample in the reference layout
 * (ms_deform_im2col_cuda.cuh:33-84 reads the four corners of a sample from four rows), whatever the element size.  msda_pack_value
 * re-lays value [N,S,M,32] (fp32 or bf16: what value_proj produces, ms_deform_attn.py:136) so that the two x-neighbours of a sample
 * share one aligned 128-byte line of bf16; msda_forward_packed then gathers two lines per sample.  Same result as msda_forward
 * with MSDA_BF16 / MSDA_BF16_LOC32 storage (value rounded to bf16, fp32 arithmetic, bf16 out).
 *   packed : msda_packed_value_bytes(N,S,M,D) bytes, 16-byte aligned; 0 = shape not supported (D must be 32)
 *   dtype of msda_pack_value     : storage type of `value` (MSDA_F32 or MSDA_BF16)
 *   dtype of msda_forward_packed : MSDA_BF16 (loc / aw bf16) or MSDA_BF16_LOC32 (loc / aw fp32); out is bf16 [N,Lq,M*32]; L*P = 16
 * shapes / level_start must be the ones the pack pass saw. */
size_t msda_packed_value_bytes(int N, int S, int M, int D);
int msda_pack_value(void* stream, int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                    int N, int S, int M, int D, int L, void* packed);
int msda_forward_packed(void* stream, int dtype, const void* packed, const int64_t* shapes, const int64_t* level_start,
                        const void* loc, const void* aw, int N, int S, int M, int D, int L, int Lq, int P, void* out);
/* The two producers / consumers that make the packed layout free for a module ("value_proj GEMM epilogue writing
 * head-major bf16 + masked_fill"; ms_deform_attn.py:136-138, :142-171):
 *   tc_linear_forward_packed : value_proj as the 3xTF32 GEMM of tc_linear_forward whose EPILOGUE writes the packed layout directly
 *     (bias added, rows with row_mask != 0 zeroed, rounded to bf16) -- no fp32 value tensor, no pack pass.  x [N*S, in_features],
 *     weight [heads*32, in_features], packed as for msda_pack_value.  Bit-identical to msda_pack_value(tc_linear_forward(x)).
 *   msda_fused_forward_packed_joint : msda_forward_packed with the fused prologue of msda_fused_forward_joint (softmax and sampling
 *     locations from the raw query projection `qproj` [N*Lq, row_stride]); out_dtype MSDA_BF16 or MSDA_F32 (what output_proj takes).
 * Inference only (the backward kernels read the reference layout). */
int tc_linear_forward_packed(void* stream, const void* x, const void* weight, const void* bias, const unsigned char* row_mask,
                             int N, int S, int in_features, int heads, const int64_t* shapes, const int64_t* level_start, int L,
                             void* packed);
int msda_fused_forward_packed_joint(void* stream, const void* packed, const int64_t* shapes, const int64_t* level_start,
                                    const void* ref_points, int R, const void* qproj, int row_stride, const void* grid, int mode,
                                    float offset_scale, int N, int S, int M, int D, int L, int Lq, int P, int out_dtype, void* out);

/* Fused sampler prologue (replaces the elementwise tail of MSDeformAttn.forward, ms_deform_attn.py:142-161):
 * the kernel takes what the module's Linear layers produce and computes softmax and sampling locations itself.
 *   offsets [N,Lq,M,L,P,2] raw output of sampling_offsets / sampling_grid_offsets,  logits [N,Lq,M,L*P] raw attention logits,
 *   ref_points [N,Lq,R] (R = 2: cx,cy; R = 4: cx,cy,w,h),  grid [M,L,P,2] = the module's `sampling_offsets` buffer (mode 1)
 *   mode 0 (pred_offsets=True, encoder):  loc = ref_xy + offsets / offset_scale
 *   mode 1 (pred_offsets=False, decoder): loc = ref_xy + (grid * 0.5 * ref_wh + clamp(offsets, +-ref_wh*offset_scale)) / offset_scale
 *   aw = softmax(logits) over L*P;  out = scale * sum_{g<G} msda(value, shapes[g], level_start[g], loc, aw)   (G = 1: spatial)
 * The backward returns d/d value, d/d offsets and d/d logits (reference points are treated as constants).
 * fp32 only, D in {32,24}, L*P in {8,16}; otherwise MSDA_ERR_UNSUPPORTED. */
int msda_fused_forward(void* stream, int dtype,
                       const void* value, const int64_t* shapes, const int64_t* level_start,
                       const void* ref_points, int R, const void* offsets, const void* logits, const void* grid,
                       int mode, float offset_scale,
                       int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                       void* out);
int msda_fused_backward(void* stream, int dtype,
                        const void* value, const int64_t* shapes, const int64_t* level_start,
                        const void* ref_points, int R, const void* offsets, const void* logits, const void* grid,
                        int mode, float offset_scale, const void* grad_out,
                        int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                        void* grad_value, void* grad_offsets, void* grad_logits);
int msda_fused_backward_flags(void* stream, int dtype,
                              const void* value, const int64_t* shapes, const int64_t* level_start,
                              const void* ref_points, int R, const void* offsets, const void* logits, const void* grid,
                              int mode, float offset_scale, const void* grad_out,
                              int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                              void* grad_value, void* grad_offsets, void* grad_logits, int flags);
/* Joint query projection: `qproj` [N*Lq, row_stride] fp32 is the output of ONE Linear layer over the concatenated weights of
 * sampling_offsets (or sampling_grid_offsets) and attention_weights (ms_deform_attn.py:143-146, :157): columns [0, 2*M*L*P) are
 * the raw offsets in [M,L,P,2] order, columns [2*M*L*P, 3*M*L*P) the raw logits in [M,L*P] order; further columns are ignored.
 * The backward writes the gradient in the same layout into grad_qproj [N*Lq, row_stride] (columns beyond 3*M*L*P are not
 * written), so the query side of a module is one GEMM forward, one for grad_query and one for the weight gradients instead of
 * two each plus an addition.  row_stride a multiple of 4; `flags` as for msda_backward_flags.  Otherwise as msda_fused_*. */
int msda_fused_forward_joint(void* stream, int dtype,
                             const void* value, const int64_t* shapes, const int64_t* level_start,
                             const void* ref_points, int R, const void* qproj, int row_stride, const void* grid,
                             int mode, float offset_scale,
                             int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                             void* out);
int msda_fused_backward_joint(void* stream, int dtype,
                              const void* value, const int64_t* shapes, const int64_t* level_start,
                              const void* ref_points, int R, const void* qproj, int row_stride, const void* grid,
                              int mode, float offset_scale, const void* grad_out,
                              int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                              void* grad_value, void* grad_qproj, int flags);

/* Mask contraction: out[b,q,n] = sum_k coeff[b,q,k] * proto[b,k,n], n over the flattened (t,h,w)
 * plane (Ncols = T*H*W).  coeff [B,Q,K], proto [B,K,Ncols], out [B,Q,Ncols].
 *   in_dtype  MSDA_F32 (3xTF32: operands split on chip into hi + lo TF32 parts, hi*hi + hi*lo + lo*hi
 *             accumulated in fp32, parity with the fp32 einsum to ~1e-6), MSDA_BF16 or MSDA_F16 (one kind::f16 pass,
 *             fp32 accumulation)
 *   out_dtype MSDA_F32, or the 16-bit type of the inputs (fp32 inputs: MSDA_F32 or MSDA_BF16)
 * Runs on the 5th-gen tensor cores (tcgen05, accumulators in TMEM). */
int mask_logits_forward(void* stream, int in_dtype, int out_dtype,
                        const void* coeff, const void* proto,
                        int B, int Q, int K, int64_t Ncols, void* out);

/* grad_coeff[b,q,k] = sum_n grad_out[b,q,n] proto[b,k,n];  grad_proto[b,k,n] = sum_q coeff[b,q,k] grad_out[b,q,n].
 * Either output may be NULL to skip it.  `dtype` must be MSDA_F32 (the reference trains the mask head in fp32).
 * Both gradients run as 3xTF32 on the tensor cores when K % 4 == 0, K <= 128, Ncols % 4 == 0 and the pointers are
 * 16-byte aligned (otherwise register-tiled SIMT kernels); grad_coeff is accumulated with atomics, so its last bits
 * depend on the schedule, like the reference's atomicAdd-based op gradients. */
int mask_logits_backward(void* stream, int dtype,
                         const void* coeff, const void* proto, const void* grad_out,
                         int B, int Q, int K, int64_t Ncols,
                         void* grad_coeff, void* grad_proto);

/* Linear layers around the sampler on the tensor cores (ms_deform_attn.py:136-138 value_proj + masked_fill,
 * :143-146 sampling offsets, :157 attention_weights, :171 output_proj; the reference runs them as fp32 cuBLAS SGEMMs):
 *   y[r, o] = sum_i x[r, i] * weight[o, i] + bias[o];  rows r with row_mask[r] != 0 are written as 0 (masked_fill)
 * x [rows, in_features], weight [out_features, in_features] (nn.Linear layout), bias [out_features] or NULL, row_mask
 * [rows] bytes or NULL, y [rows, out_features]; all fp32, device memory.  3xTF32 (hi*hi + hi*lo + lo*hi, fp32
 * accumulation): ~1e-6 of the exact fp32 product.  in_features and out_features must be multiples of 4.
 * tc_linear_backward: grad_x[r, i] = sum_o grad_y[r, o] weight[o, i]; grad_weight[o, i] = sum_r grad_y[r, o] x[r, i]
 * (zero-filled inside, accumulated across CTAs with TMA reduce-add stores); either output may be NULL. */
int tc_linear_forward(void* stream, const void* x, const void* weight, const void* bias, const unsigned char* row_mask,
                      int64_t rows, int in_features, int out_features, void* y);
int tc_linear_backward(void* stream, const void* grad_y, const void* x, const void* weight,
                       int64_t rows, int in_features, int out_features, void* grad_x, void* grad_weight);
/* grad_bias[o] = sum_r grad_y[r, o] (nn.Linear's bias gradient; torch's strided sum(0) costs 28 us per encoder-sized call, this
 * streams grad_y once at memory speed).  grad_bias is zero-filled inside; out_features a multiple of 4. */
int tc_linear_bias_grad(void* stream, const void* grad_y, int64_t rows, int out_features, void* grad_bias);
/* tc_linear_backward with the bias gradient in the same call (grad_bias [out_features] or NULL, zero-filled inside).  When
 * grad_weight is requested the column sums of grad_y are taken inside the weight-gradient GEMM by the warps that stage grad_y
 * for the tensor cores (no extra pass over grad_y, no extra launch; accumulated with atomics across the CTAs of the split
 * reduction); otherwise this is tc_linear_backward followed by tc_linear_bias_grad. */
int tc_linear_backward_bias(void* stream, const void* grad_y, const void* x, const void* weight,
                            int64_t rows, int in_features, int out_features, void* grad_x, void* grad_weight, void* grad_bias);

/* Host-buffer entries: same semantics, every pointer is HOST memory (pinned memory makes the
 * copies asynchronous).  The library owns a grow-only device arena per process; `device` selects
 * the GPU.  These return after the results have landed in the host output buffers. */
int msda_forward_host(int device, int dtype,
                      const void* value, const int64_t* shapes, const int64_t* level_start,
                      const void* loc, const void* aw,
                      int N, int S, int M, int D, int L, int Lq, int P,
                      void* out);
int msda_backward_host(int device, int dtype,
                       const void* value, const int64_t* shapes, const int64_t* level_start,
                       const void* loc, const void* aw, const void* grad_out,
                       int N, int S, int M, int D, int L, int Lq, int P,
                       void* grad_value, void* grad_loc, void* grad_aw);
int mask_logits_forward_host(int device, int in_dtype, int out_dtype,
                             const void* coeff, const void* proto,
                             int B, int Q, int K, int64_t Ncols, void* out);
/* With msda_set_option("host_async", 1) the *_host entries return as soon as their copies and kernels are enqueued
 * (upload, kernels and download of consecutive calls overlap on three streams; PCIe runs full duplex); host output
 * buffers are valid, and host input buffers may be reused, only after msda_host_sync().  Returns the first error. */
int msda_host_sync(void);
/* Finer-grained completion for callers that keep several steps in flight (a data-parallel trainer prefetching the next
 * clip: the uploads of step i+1's forward then run under the downloads of step i's backward and PCIe stays full duplex
 * across step boundaries).  msda_host_fence() marks everything enqueued so far and returns a ticket;
 * msda_host_wait(ticket) returns once all of that work's results are in host memory (its host input buffers may be
 * reused, its staging space in the device arena is recycled).  Tickets complete in order; ticket 0 is always complete;
 * msda_host_sync() completes every outstanding ticket. */
int msda_host_fence(int64_t* ticket);
int msda_host_wait(int64_t ticket);

/* "Saved" host entries: the host-buffer counterpart of autograd's save_for_backward (ms_deform_attn_func.py:28-29 saves
 * value, shapes, level_start_index, loc and aw on the device between forward and backward).  The forward keeps the device
 * copies of its inputs in a pooled block and returns a handle in *saved; the backward takes the handle, uploads only
 * grad_out and consumes the handle.  A forward whose backward never runs (inference) must call msda_host_saved_release().
 * G / scale as in msda_forward_grouped (G = 1, scale = 1: the plain operator; shapes [G,L,2], level_start [G,L]).
 * Halves the host->device traffic of a training step and lets the clip-level attention go up as ONE value tensor. */
int msda_forward_host_saved(int device, int dtype,
                            const void* value, const int64_t* shapes, const int64_t* level_start,
                            const void* loc, const void* aw,
                            int N, int S, int M, int D, int G, int L, int Lq, int P, float scale,
                            void* out, int64_t* saved);
int msda_backward_host_saved(int64_t saved, const void* grad_out,
                             void* grad_value, void* grad_loc, void* grad_aw);
int mask_logits_forward_host_saved(int device, int in_dtype, int out_dtype,
                                   const void* coeff, const void* proto,
                                   int B, int Q, int K, int64_t Ncols, void* out, int64_t* saved);
int mask_logits_backward_host_saved(int64_t saved, const void* grad_out, void* grad_coeff, void* grad_proto);
int msda_host_saved_release(int64_t saved);
/* Release the device arena used by the *_host entries. */
int msda_host_arena_release(void);

/* ---- callers either side of the path; fp32, device pointers, row-major contiguous --------------------
 *
 * mask_match_cost: the Hungarian matcher's mask costs (mdqe/models/matcher.py:182-200: out_masks = einsum(mask_coeff, proto);
 * cost_bce = batch_sigmoid_ce_loss(out_masks, tgt) (:36-61); cost_dice = batch_dice_loss(out_masks, tgt) (:11-28)) for ONE clip,
 * fused with the contraction: coeff [Q,K] (K <= 32), proto [K,Ncols], targets [G,Ncols] -> cost_bce [Q,G], cost_dice [Q,G].
 * out_masks is never materialised.  workspace: mask_match_cost_workspace_bytes() bytes of device memory. */
size_t mask_match_cost_workspace_bytes(void);
int mask_match_cost(void* stream, const void* coeff, const void* proto, const void* targets,
                    int Q, int K, int G, int64_t Ncols, void* workspace, void* cost_bce, void* cost_dice);
/* mask_losses_*: the criterion's mask losses of the G matched queries of a batch (mdqe/models/criterion.py:440-473): with
 * src_masks = coeff[G,K] . proto[K,Ncols] (never materialised), loss_mask / loss_dice are sigmoid_ce_loss / dice_loss (:87-108, :20-43)
 * when targets_interinst is NULL and interinst_sigmoid_ce_loss / interinst_dice_loss (:116-145, :51-81) otherwise
 * (targets, targets_interinst [G,Ncols]; all rows must share one proto, i.e. one call per clip).  G <= 32 per call, K <= 32.
 * forward:  losses[2] = (loss_mask, loss_dice), row_stats [G,8] = per-row plane sums kept for the backward.
 * backward: grad_losses[2] on the device (upstream gradients of the two losses) -> grad_coeff [G,K], grad_proto [K,Ncols]. */
size_t mask_losses_workspace_bytes(void);
int mask_losses_forward(void* stream, const void* coeff, const void* proto, const void* targets, const void* targets_interinst,
                        int G, int K, int64_t Ncols, float num_masks, void* workspace, void* row_stats, void* losses);
int mask_losses_backward(void* stream, const void* coeff, const void* proto, const void* targets, const void* targets_interinst,
                         const void* row_stats, const void* grad_losses, int G, int K, int64_t Ncols, float num_masks,
                         void* grad_coeff, void* grad_proto);
/* mask_nms_siou: soft-IoU matrix of inference_clip (mdqe/mdqe.py:394-401): mask_pred [Q,T,H,W] -> siou [Q,Q] with
 * mask_nms = mask_pred[:, ::2] if T >= 5, nearest 0.5x downsampling, soft = sigmoid, hard = soft > 0.5,
 * siou = soft.hard^T / (sum soft [:,None] + sum hard [None] - soft.hard^T + 1). */
size_t mask_nms_siou_workspace_bytes(void);
int mask_nms_siou(void* stream, const void* mask_pred, int Q, int T, int H, int W, void* workspace, void* siou);
/* mask_track_siou: the tracker's mask IoU (mdqe/tracking/OverTracker.py:92-113, OverTracker._get_siou): saved_masks [Ns,T,H,W] and
 * input_masks [Ni,T,H,W] are probabilities, thresholded at 0.5; siou [Ns,Ni] = |s & i| / (|s| + |i| - |s & i| + 1e-6), 0 for pairs
 * with an empty mask.  workspace: mask_nms_siou_workspace_bytes(). */
int mask_track_siou(void* stream, const void* saved_masks, const void* input_masks, int Ns, int Ni, int T, int H, int W,
                    void* workspace, void* siou);
/* aligned_bilinear (mdqe/util/misc.py:485-507) of n_img planes [H,W] by an integer factor, optionally followed by the sigmoid of
 * mdqe/mdqe.py:357: out [n_img, factor*H, factor*W]. */
int aligned_bilinear_sigmoid(void* stream, const void* in, int64_t n_img, int H, int W, int factor, int apply_sigmoid, void* out);
/* Query initialisation sampling (mdqe/models/transformer_dec.py:170-179): for every level l, F.grid_sample(feat_l, 2*coords-1,
 * bilinear, padding_mode="border", align_corners=False), mean over the L levels.  feat [B,S,C] (C % 4 == 0), shapes [L,2] and
 * level_start [L] int64 on the device, coords [B,Q,2] normalised (x, y) -> out [B,Q,C].  The backward zero-fills grad_feat
 * [B,S,C] itself and returns grad_coords [B,Q,2]. */
int query_init_sample_forward(void* stream, const void* feat, const int64_t* shapes, const int64_t* level_start,
                              const void* coords, int B, int S, int C, int L, int Q, void* out);
int query_init_sample_backward(void* stream, const void* feat, const int64_t* shapes, const int64_t* level_start,
                               const void* coords, const void* grad_out, int B, int S, int C, int L, int Q,
                               void* grad_feat, void* grad_coords);

/* Per-launch kernel timing (CUDA events on the launching stream, recorded right around the kernel).
 * Enable with msda_set_option("profile", 1); every launch of the given kind since the last read whose
 * work size (pairs N*Lq*M for MSDA, B*Q*Ncols for the mask kernels) is >= min_units is summed into
 * *total_ms and counted in *count; the records of that kind are then dropped.  Synchronises on the
 * recorded events.  Launches made under CUDA-graph capture are not recorded. */
#define MSDA_PROF_MSDA_FWD 0
#define MSDA_PROF_MSDA_BWD 1
#define MSDA_PROF_MASK_FWD 2
#define MSDA_PROF_MASK_BWD 3
int msda_profile_read(int kind, int64_t min_units, double* total_ms, int64_t* count);

/* Debugging aid for the pipelined tensor-core mask kernel: with option "mask_debug" = 1 CTA 0 records clock64()
 * stamps for its first 16 work items; this copies the 5 x 16 table (rows: TMA issued, MMA thread starts waiting,
 * operands landed, epilogue starts, epilogue done) to `host80` (80 values). */
int msda_debug_read(long long* host80);

/* Number of kernels this library has launched since load / since the last reset (bench.py's
 * gpu_launches claim is read from here). */
int64_t msda_launch_count(void);
void msda_launch_count_reset(void);
/* The Linear-layer GEMMs combine tiles that are split between CTAs through per-tile flags with BOUNDED waits (csrc/gemm3x.cuh): a wait
 * that expires is counted here instead of hanging the GPU.  Synchronises the current device; 0 on a healthy run, -1 without a device. */
int64_t msda_gemm_flag_timeouts(void);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
