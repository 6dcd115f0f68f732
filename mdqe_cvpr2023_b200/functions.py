"""Autograd surface of the hot path.

`MSDeformAttnFunction` keeps the reference signature
``apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step)``
(/root/reference/mdqe/models/ops/functions/ms_deform_attn_func.py:22-42): gradients for arguments
0, 3 and 4 only, once-differentiable, inputs cast to fp32 under autocast (custom_fwd(cast_inputs=float32)),
so the default behaviour equals the reference.  bf16 is opt-in: tensors that arrive as bf16 OUTSIDE
autocast run the bf16 kernels (fp32 arithmetic).

`mask_logits` is the operator form of the four ``torch.einsum('bqm,bmthw->bqthw')`` sites
(transformer_dec.py:255, mdqe/mdqe.py:384, matcher.py:182, criterion.py:440).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops

# The backward accumulates grad_value with reductions into a zero-filled tensor.  Filling 20.9 MB right before the kernel (what
# the reference's `zeros_like` does, ms_deform_attn_cuda.cu:121) costs 7 us per call on the critical path; here the forward
# allocates and zero-fills that tensor on a side stream, where it overlaps the forward kernels, and the backward only waits
# for the fill's event (a B200 has the memory to keep the buffers from forward to backward: 0.75 GB per R50_ovis_360 clip).
# Set to False to allocate and fill inside the backward like the reference.
PREZERO_GRAD_VALUE = True
_side_streams = {}


def _prezero(ctx, value):
    ctx.acc = ctx.acc_ready = None
    if not (PREZERO_GRAD_VALUE and ctx.needs_input_grad[0]):
        return
    dev = value.device
    cur = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(dev)
    if torch.cuda.is_current_stream_capturing():
        side.wait_stream(cur)                       # fork inside the capture; the backward's wait_event is the join
    with torch.cuda.stream(side):
        ctx.acc = ops.new_backward_accumulator(value)
        ctx.acc_ready = side.record_event()


def _take_accumulator(ctx, device):
    acc = ctx.acc
    if acc is None:
        return None
    cur = torch.cuda.current_stream(device)
    cur.wait_event(ctx.acc_ready)
    acc.record_stream(cur)
    ctx.acc = ctx.acc_ready = None
    return acc


class MSDeformAttnFunction(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ops.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                            attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        _prezero(ctx, value)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        value, shapes, level_start, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = ops.ms_deform_attn_backward(
            value, shapes, level_start, loc, aw, grad_output.contiguous(), ctx.im2col_step, _take_accumulator(ctx, value.device))
        return grad_value, None, None, grad_loc, grad_aw, None


class MSDeformAttnGroupedFunction(Function):
    """apply(value, spatial_shapes[G,L,2], level_start_index[G,L], sampling_locations, attention_weights, scale):
    scale * sum over the G level tables, one kernel launch forward and one backward.  This is the clip-level ("temporal")
    attention of the reference module (ms_deform_attn.py:219-235: one Function call per pyramid level, then a mean)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, spatial_shapes, level_start_index, sampling_locations, attention_weights, scale):
        ctx.scale = float(scale)
        output = ops.ms_deform_attn_grouped_forward(value, spatial_shapes, level_start_index, sampling_locations,
                                                    attention_weights, ctx.scale)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
        _prezero(ctx, value)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        value, shapes, level_start, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = ops.ms_deform_attn_grouped_backward(
            value, shapes, level_start, loc, aw, grad_output.contiguous(), ctx.scale, _take_accumulator(ctx, value.device))
        return grad_value, None, None, grad_loc, grad_aw, None


class MSDeformAttnFusedFunction(Function):
    """apply(value, spatial_shapes, level_start_index, reference_points, offsets, logits, grid, mode, offset_scale, scale):
    the sampler with softmax and sampling-location arithmetic inside the kernel (ms_deform_attn.py:142-161 fused, SURVEY 8f
    N1).  `offsets` / `logits` are the raw outputs of the module's Linear layers; gradients flow to value, offsets, logits
    (reference points are constants on this path)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, offsets, logits, grid, mode, offset_scale, scale):
        ctx.cfg = (int(mode), float(offset_scale), float(scale))
        out = ops.ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, offsets, logits,
                                               grid, *ctx.cfg)
        ctx.has_grid = grid is not None
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, offsets, logits,
                              *([grid] if grid is not None else []))
        _prezero(ctx, value)
        return out

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        saved = ctx.saved_tensors
        value, shapes, starts, ref, offsets, logits = saved[:6]
        grid = saved[6] if ctx.has_grid else None
        mode, offset_scale, scale = ctx.cfg
        gv, goff, glog = ops.ms_deform_attn_fused_backward(value, shapes, starts, ref, offsets, logits, grid, mode, offset_scale,
                                                           grad_output.contiguous(), scale, _take_accumulator(ctx, value.device))
        return gv, None, None, None, goff, glog, None, None, None, None


class MSDeformAttnFusedJointFunction(Function):
    """apply(value, spatial_shapes, level_start_index, reference_points, qproj, n_points, grid, mode, offset_scale, scale):
    MSDeformAttnFusedFunction with the raw offsets and logits given as column ranges of one tensor ``qproj`` [N,Lq,3*M*L*P]
    (offsets first) -- the output of ONE Linear layer over the concatenated sampling_offsets / attention_weights parameters.
    The gradient comes back in the same layout, so the query side of a module costs one GEMM forward, one for grad_query and
    one for the weight gradients (instead of two each and an addition)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, qproj, n_points, grid, mode, offset_scale, scale):
        ctx.cfg = (int(n_points), int(mode), float(offset_scale), float(scale))
        n_points, mode, offset_scale, scale = ctx.cfg
        out = ops.ms_deform_attn_fused_forward_joint(value, spatial_shapes, level_start_index, reference_points, qproj, n_points,
                                                     grid, mode, offset_scale, scale)
        ctx.has_grid = grid is not None
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, qproj, *([grid] if grid is not None else []))
        _prezero(ctx, value)
        return out

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        saved = ctx.saved_tensors
        value, shapes, starts, ref, qproj = saved[:5]
        grid = saved[5] if ctx.has_grid else None
        n_points, mode, offset_scale, scale = ctx.cfg
        gv, gq = ops.ms_deform_attn_fused_backward_joint(value, shapes, starts, ref, qproj, n_points, grid, mode, offset_scale,
                                                         grad_output.contiguous(), scale, _take_accumulator(ctx, value.device))
        return gv, None, None, None, gq, None, None, None, None, None


class _MaskLogitsFunction(Function):
    @staticmethod
    def forward(ctx, coeff, proto):
        ctx.save_for_backward(coeff, proto)
        return ops.mask_logits_forward(coeff, proto)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        coeff, proto = ctx.saved_tensors
        gc, gp = ops.mask_logits_backward(coeff, proto, grad_out.contiguous(),
                                          need_coeff=ctx.needs_input_grad[0], need_proto=ctx.needs_input_grad[1])
        return gc, gp


def mask_logits(coeff, proto):
    """Drop-in for ``torch.einsum('bqm,bmthw->bqthw', coeff, proto)`` (also accepts the unbatched
    'qm,mthw->qthw' form of mdqe/mdqe.py:384).

    Dtypes follow the einsum it replaces.  Under ``torch.autocast`` -- how the reference evaluates (train_net.py:207-208; neither
    Transformer_Dec.forward nor inference_clip disables it, so mdqe/mdqe.py:384 and transformer_dec.py:255 run in fp16) -- both
    operands are cast to the autocast dtype and the result has that dtype; outside autocast float32, bfloat16 and float16
    operands of one common dtype give a result of that dtype (mixed dtypes raise, like einsum)."""
    if torch.is_autocast_enabled("cuda") and coeff.is_cuda:
        dt = torch.get_autocast_dtype("cuda")
        coeff, proto = coeff.to(dt), proto.to(dt)
        with torch.autocast("cuda", enabled=False):
            return mask_logits(coeff, proto)
    if coeff.dim() == 2:
        return _MaskLogitsFunction.apply(coeff.unsqueeze(0).contiguous(), proto.unsqueeze(0).contiguous()).squeeze(0)
    return _MaskLogitsFunction.apply(coeff.contiguous(), proto.contiguous())


class _TcLinearFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, row_mask):
        ctx.save_for_backward(x, weight, row_mask)
        ctx.has_bias = bias is not None
        return ops.tc_linear_forward(x, weight, bias, row_mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x, weight, row_mask = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        if row_mask is not None:
            grad_y = grad_y.masked_fill(row_mask.unsqueeze(-1), 0.0)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gx, gw, gb = ops.tc_linear_backward(grad_y, x, weight, need_x=ctx.needs_input_grad[0], need_weight=ctx.needs_input_grad[1], need_bias=True)
        else:
            (gx, gw), gb = ops.tc_linear_backward(grad_y, x, weight, need_x=ctx.needs_input_grad[0], need_weight=ctx.needs_input_grad[1]), None
        return gx, gw, gb, None


def tc_linear(x, weight, bias=None, row_mask=None):
    """``F.linear(x, weight, bias)`` (optionally followed by ``masked_fill(row_mask[..., None], 0)``) on the tensor cores in
    3xTF32 -- fp32-level accuracy without the fp32 SIMT GEMM (ms_deform_attn.py:136-138, :143-146, :157, :171)."""
    return _TcLinearFunction.apply(x.contiguous(), weight.contiguous(), bias, row_mask.contiguous() if row_mask is not None else None)
