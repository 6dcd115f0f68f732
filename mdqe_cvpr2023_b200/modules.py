"""`MSDeformAttn` with the reference module's constructor, forward signature and state-dict keys
(/root/reference/mdqe/models/ops/modules/ms_deform_attn.py:34-238), running on the B200 kernels.

Kept identical so that mdqe/models/transformer_{enc,dec}.py and released checkpoints work as-is:
  * parameters  value_proj, output_proj, attention_weights, and sampling_offsets (pred_offsets=True,
    encoder) or sampling_grid_offsets (pred_offsets=False, decoder)            (ms_deform_attn.py:68-74)
  * buffers     lvl_spatial_scales, and sampling_offsets when pred_offsets=False  (:63-66, :96)
  * `_reset_parameters()` is public in practice: the decoder calls it again    (transformer_dec.py:72-74)
  * mode 'spatial' samples the L pyramid levels of one frame; mode 'temporal' samples the T frames of
    the clip level by level and averages the levels                           (:118-173, :175-238)

What differs from the reference implementation (not from its results):
  * temporal mode hands the operator a *view* of the whole [B, T*S, M, D] value tensor and encodes
    "frame t of level l" in level_start_index (= t*S + start_l) instead of materialising four
    `.contiguous()` copies per call (ms_deform_attn.py:222-224); on the fast-kernel configurations all
    pyramid levels run in ONE launch (MSDeformAttnGroupedFunction) with the mean folded in;
  * all tensor work below the Linear layers goes through MSDeformAttnFunction -> libmsda_b200.so;
  * the four Linear layers themselves (value_proj + masked_fill, offsets, attention weights, output_proj) run as 3xTF32
    tensor-core GEMMs (`tc_linear`, csrc/gemm3x.cuh) when `self.tc_linear` is set (default) and the tensors are fp32 on CUDA
    -- same parameters, fp32-level accuracy, 2.7-4x less GPU time than the fp32 SGEMMs.
"""
import math
import warnings

import weakref

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .functions import (MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnFusedJointFunction, MSDeformAttnGroupedFunction,
                        tc_linear)


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, n_frames=1, pred_offsets=True,
                 mode='spatial'):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f'd_model must be divisible by n_heads, but got {d_model} and {n_heads}')
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-2 head dimension is the fast path of the reference CUDA "
                          "kernels; here 32 and 24 are specialised and every other size takes the generic kernel.")
        if mode not in ('spatial', 'temporal'):
            raise ValueError(f"mode must be 'spatial' or 'temporal', got {mode!r}")

        self.im2col_step = 64
        self.mode = mode
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.pred_offsets = pred_offsets
        self.scale = 8.
        self.n_frames = n_frames
        # fold softmax + sampling-location arithmetic into the sampler kernel where it is implemented (fp32, D in
        # {32,24}, L*P in {8,16}, reference points without gradient); set False to run the reference's op sequence
        self.fused_prologue = True
        self.tc_linear = True
        self.joint_query_proj = True       # fused prologue + tc_linear: offsets and logits from ONE GEMM over the concatenated weights
        # "bf16_packed": inference (no autograd graph) in spatial mode keeps `value` only as the sampler's paired-corner bf16 layout,
        # written by the value_proj GEMM epilogue (north_star's bf16 mode; value rounded to bf16, everything else fp32: <= 2e-2)
        self.value_storage = "fp32"
        self.value_storage_always = False  # True: also for calls with few queries (tests / A-B)
        self.check_shapes = False          # True: assert sum(H_l * W_l) == S like the reference (one host sync per call)

        # what the operator sees as "levels": pyramid levels of a frame, or the frames of a clip
        if mode == 'spatial':
            self.lvl = n_levels
            lvl_spatial_scales = torch.arange(1, self.lvl + 1)
        else:
            self.lvl = n_frames
            lvl_spatial_scales = torch.full((self.lvl,), 2, dtype=torch.long)
        self.register_buffer("lvl_spatial_scales", lvl_spatial_scales)

        n_samples = n_heads * self.lvl * n_points
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.attention_weights = nn.Linear(d_model, n_samples)
        if pred_offsets:
            self.sampling_offsets = nn.Linear(d_model, n_samples * 2)
        else:
            self.sampling_grid_offsets = nn.Linear(d_model, n_samples * 2)
        self._reset_parameters()

    # ------------------------------------------------------------------------------------- init
    def _reset_parameters(self):
        H, L, K = self.n_heads, self.lvl, self.n_points
        angle = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
        ray = torch.stack([angle.cos(), angle.sin()], -1)
        ray = ray / ray.abs().max(-1, keepdim=True)[0]                       # unit square directions, one per head
        steps = torch.arange(1, K + 1, dtype=torch.float32).view(1, 1, K, 1)
        grid = ray.view(H, 1, 1, 2).repeat(1, L, K, 1) * steps               # H L K 2: k-th point k+1 steps out
        grid = grid / K * self.scale
        if self.pred_offsets:
            nn.init.constant_(self.sampling_offsets.weight.data, 0.)
            grid = grid * 0.05 * self.lvl_spatial_scales.reshape(1, -1, 1, 1).to(grid)
            with torch.no_grad():
                self.sampling_offsets.bias = nn.Parameter(grid.reshape(-1))
        else:
            self.register_buffer("sampling_offsets", grid.view(1, 1, H, L, K, 2).clone())
            nn.init.constant_(self.sampling_grid_offsets.weight.data, 0.)
            nn.init.constant_(self.sampling_grid_offsets.bias.data, 0.)
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)

    # ---------------------------------------------------------------------------------- pieces
    def _linear(self, layer, x, row_mask=None):
        """layer(x) (+ masked_fill of the rows selected by row_mask) -- on the tensor cores when possible."""
        if self.tc_linear and ops.linear_supported(x, layer.weight) and not torch.is_autocast_enabled():
            return tc_linear(x, layer.weight, layer.bias, row_mask)
        y = layer(x)
        if row_mask is not None:
            y = y.masked_fill(row_mask[..., None], float(0))
        return y

    def _sampling(self, query, reference_points):
        """-> (sampling_locations [B,Q,H,L,K,2], attention_weights [B,Q,H,L,K]) for L = self.lvl."""
        B, Q, _ = query.shape
        H, L, K = self.n_heads, self.lvl, self.n_points
        ref = reference_points.view(B, Q, 1, 1, 1, -1)
        if self.pred_offsets:
            offsets = self._linear(self.sampling_offsets, query).view(B, Q, H, L, K, 2)
        else:
            # fixed ray grid scaled to half the reference box, plus a learned residual clamped to +-8 boxes
            box = ref[..., 2:]
            residual = self._linear(self.sampling_grid_offsets, query).view(B, Q, H, L, K, 2).to(ref)
            bound = box * self.scale
            residual = torch.where(residual > -bound, residual, -bound)
            residual = torch.where(residual < bound, residual, bound)
            offsets = self.sampling_offsets * 0.5 * box + residual
        locations = ref[..., :2] + offsets / self.scale
        logits = self._linear(self.attention_weights, query).view(B, Q, H, L * K)
        weights = F.softmax(logits, -1).view(B, Q, H, L, K)
        return locations, weights

    def _fused_inputs(self, query):
        """raw Linear outputs in the layouts msda_fused_forward expects (views, no copies)."""
        B, Q, _ = query.shape
        H, L, K = self.n_heads, self.lvl, self.n_points
        lin = self.sampling_offsets if self.pred_offsets else self.sampling_grid_offsets
        offsets = self._linear(lin, query).view(B, Q, H, L, K, 2)
        logits = self._linear(self.attention_weights, query).view(B, Q, H, L * K)
        grid = None if self.pred_offsets else self.sampling_offsets.reshape(H, L, K, 2).contiguous()
        return offsets, logits, grid, (0 if self.pred_offsets else 1)

    def _joint_query_projection(self, query):
        """sampling_offsets (or sampling_grid_offsets) and attention_weights as ONE Linear layer over the concatenated parameters
        -> [B,Q,3*H*L*K] with the raw offsets in the first 2*H*L*K columns; None where the tensor-core Linear does not apply.
        The two `cat`s are the only extra work; autograd hands their gradient back to the two layers as views."""
        lin = self.sampling_offsets if self.pred_offsets else self.sampling_grid_offsets
        aw = self.attention_weights
        if not (self.joint_query_proj and self.tc_linear and ops.linear_supported(query, lin.weight)) or torch.is_autocast_enabled():
            return None
        weight = torch.cat([lin.weight, aw.weight], 0)
        bias = torch.cat([lin.bias, aw.bias], 0)
        return tc_linear(query, weight, bias)

    def _fused_sampler(self, value, shapes, starts, reference_points, query, scale):
        reference_points = reference_points.contiguous()
        qproj = self._joint_query_projection(query)
        if qproj is not None:
            grid = None if self.pred_offsets else self.sampling_offsets.reshape(self.n_heads, self.lvl, self.n_points, 2).contiguous()
            return MSDeformAttnFusedJointFunction.apply(value, shapes, starts, reference_points, qproj, self.n_points, grid,
                                                        0 if self.pred_offsets else 1, self.scale, scale)
        offsets, logits, grid, mode = self._fused_inputs(query)
        return MSDeformAttnFusedFunction.apply(value, shapes, starts, reference_points, offsets, logits, grid, mode, self.scale, scale)

    def _packed_inference_ok(self, query, reference_points, input_flatten):
        B, S, _ = input_flatten.shape
        return (not (torch.is_grad_enabled() and (query.requires_grad or input_flatten.requires_grad
                                                  or any(p.requires_grad for p in self.parameters())))
                and self.fused_prologue and self.joint_query_proj and self.tc_linear and not torch.is_autocast_enabled()
                and input_flatten.is_cuda and input_flatten.dtype == torch.float32 and query.dtype == torch.float32
                and self.d_model == self.n_heads * 32 and self.lvl * self.n_points == 16 and self.lvl <= 32
                and reference_points.dtype == torch.float32 and reference_points.shape[-1] in (2, 4)
                and ops.linear_supported(input_flatten, self.value_proj.weight) and ops.linear_supported(query, self.attention_weights.weight)
                and B * 2 * S * self.n_heads < 2 ** 32 and B * query.shape[1] * self.n_heads < 2 ** 31 // 64
                # pays where the sampler dominates (encoder: one query per pixel: -16 us per call); with 196 decoder queries the
                # sampler is launch-bound either way and the packed epilogue only costs (profiles/r02av_packed_module.txt)
                and (query.shape[1] * 2 >= S or self.value_storage_always))

    def _packed_inference(self, query, reference_points, input_flatten, shapes_c, level_start, input_padding_mask):
        """value_proj -> packed bf16 value (GEMM epilogue), ONE query-projection GEMM, packed sampler with the fused prologue, output_proj"""
        B, S, _ = input_flatten.shape
        packed = ops.tc_linear_forward_packed(input_flatten.contiguous(), self.value_proj.weight, self.value_proj.bias,
                                              input_padding_mask.contiguous() if input_padding_mask is not None else None,
                                              shapes_c, level_start, self.n_heads)
        qproj = self._joint_query_projection(query)
        grid = None if self.pred_offsets else self.sampling_offsets.reshape(self.n_heads, self.lvl, self.n_points, 2).contiguous()
        sampled = ops.ms_deform_attn_fused_forward_packed_joint(packed, (B, S, self.n_heads, 32), shapes_c, level_start,
                                                                reference_points.contiguous(), qproj, self.n_points, grid,
                                                                0 if self.pred_offsets else 1, self.scale, torch.float32)
        return self._linear(self.output_proj, sampled)

    _geometry_cache = []

    def _project_value(self, input_flatten, input_padding_mask):
        value = self._linear(self.value_proj, input_flatten, input_padding_mask)
        return value.view(*value.shape[:-1], self.n_heads, self.d_model // self.n_heads)

    @staticmethod
    def _geometry(spatial_shapes, n_frames=0, rows_per_frame=0):
        """level_start_index and (temporal mode) the [G,T,2] / [G,T] level tables, computed on the device without a host
        sync.  The reference asserts sum(H_l*W_l) == S on the host (ms_deform_attn.py:134: one sync per call); here the kernels
        check every level window against the S rows they may touch and disable a level that does not fit
        (csrc/msda_common.cuh level_fits) -- a mismatch gives zeros for that level, never an out-of-bounds access.  Set
        MSDeformAttn.check_shapes = True to get the reference's assertion (and its sync) back.  Callers rebuild
        `spatial_shapes` every forward (transformer_enc.py:46) and multi-scale training changes its contents, so nothing is
        remembered by value -- but the 18 attention modules of one forward pass all receive the SAME tensor object, and the
        tables are shared between them (keyed on the object's identity and version counter: ~8 tiny kernels per module call
        otherwise, 0.2 ms of the module-level step)."""
        for ref, ver, nf, rpf, out in MSDeformAttn._geometry_cache:
            if ref() is spatial_shapes and ver == spatial_shapes._version and nf == n_frames and rpf == rows_per_frame:
                return out
        sizes = spatial_shapes.prod(-1)
        starts = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]]).long()
        shapes_c = spatial_shapes.contiguous()
        shapes_g = starts_g = None
        if n_frames:
            n_lvl = spatial_shapes.shape[0]
            frame_base = torch.arange(n_frames, device=starts.device, dtype=starts.dtype) * rows_per_frame
            shapes_g = shapes_c.view(n_lvl, 1, 2).expand(n_lvl, n_frames, 2).contiguous()
            starts_g = (starts.view(n_lvl, 1) + frame_base.view(1, n_frames)).contiguous()
        out = (starts, shapes_c, shapes_g, starts_g)
        cache = MSDeformAttn._geometry_cache
        cache.append((weakref.ref(spatial_shapes), spatial_shapes._version, n_frames, rows_per_frame, out))
        del cache[:-4]                                          # spatial + temporal tables of the current and the previous forward
        return out

    # --------------------------------------------------------------------------------- forward
    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_padding_mask=None):
        if self.mode == 'spatial':
            return self.spatial_forward(query, reference_points, input_flatten, input_spatial_shapes,
                                        input_padding_mask)
        return self.temporal_clip_forward(query, reference_points, input_flatten, input_spatial_shapes,
                                          input_padding_mask)

    @torch.amp.autocast("cuda", enabled=False)
    def spatial_forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_padding_mask=None):
        """query BxQxC, reference_points BxQx4 (cx, cy, w, h), input_flatten BxSxC with S = sum_l H_l*W_l."""
        if self.check_shapes:
            assert int(input_spatial_shapes.prod(-1).sum()) == input_flatten.shape[1], "spatial_shapes do not add up to the rows of input_flatten"
        level_start, shapes_c, _, _ = self._geometry(input_spatial_shapes)
        if self.value_storage == "bf16_packed" and self._packed_inference_ok(query, reference_points, input_flatten):
            return self._packed_inference(query, reference_points, input_flatten, shapes_c, level_start, input_padding_mask)
        value = self._project_value(input_flatten, input_padding_mask).contiguous()  # B S H D
        if self.fused_prologue and ops.fused_supported(value, reference_points, 1, self.lvl, self.n_points, query.shape[1]):
            sampled = self._fused_sampler(value, shapes_c, level_start, reference_points, query, 1.0)
            return self._linear(self.output_proj, sampled)
        locations, weights = self._sampling(query, reference_points)
        sampled = MSDeformAttnFunction.apply(value.contiguous(), shapes_c, level_start,
                                             locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self._linear(self.output_proj, sampled)

    @torch.amp.autocast("cuda", enabled=False)
    def temporal_clip_forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                              input_padding_mask=None):
        """query BxQxC, reference_points BxQx4, input_flatten BxTxSxC; the T frames play the role of levels."""
        B, T, S, _ = input_flatten.shape
        assert T == self.n_frames, f"temporal MSDeformAttn built for {self.n_frames} frames, got {T}"
        if self.check_shapes:
            assert int(input_spatial_shapes.prod(-1).sum()) == S, "spatial_shapes do not add up to the rows of input_flatten"
        level_start, _, shapes_g, starts_g = self._geometry(input_spatial_shapes, T, S)
        value = self._project_value(input_flatten, input_padding_mask)               # B T S H D
        value = value.contiguous().view(B, T * S, self.n_heads, -1)                  # frames back to back, no copies
        n_lvl = input_spatial_shapes.shape[0]
        if self.fused_prologue and ops.fused_supported(value, reference_points, n_lvl, T, self.n_points, query.shape[1]):
            # one launch: all pyramid levels (grouped form) + softmax / location arithmetic inside the kernel
            sampled = self._fused_sampler(value, shapes_g, starts_g, reference_points, query, 1.0 / n_lvl)
            return self._linear(self.output_proj, sampled)
        locations, weights = self._sampling(query, reference_points)
        locations, weights = locations.contiguous(), weights.contiguous()
        if ops.grouped_supported(value, n_lvl, T, self.n_points, query.shape[1], locations, weights):
            # all pyramid levels in ONE launch: level table g = the T frames of pyramid level g, mean folded in
            sampled = MSDeformAttnGroupedFunction.apply(value, shapes_g, starts_g, locations, weights, 1.0 / n_lvl)
            return self._linear(self.output_proj, sampled)
        sampled = None
        for lvl in range(n_lvl):
            shapes_l, starts_l = shapes_g[lvl], starts_g[lvl]
            out_l = MSDeformAttnFunction.apply(value, shapes_l, starts_l, locations, weights, self.im2col_step)
            sampled = out_l if sampled is None else sampled + out_l
        sampled = sampled / input_spatial_shapes.shape[0]
        return self._linear(self.output_proj, sampled)
