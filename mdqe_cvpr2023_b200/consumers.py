"""Host side of the callers either side of the hot path: the functions the reference's matcher,
inference head and query initialisation would call instead of their chains of torch ops.  CUDA tensors only -- like the
operator itself there is no CPU or PyTorch fallback (CPU tensors raise RuntimeError).

  mask_match_cost(mask_coeff, proto, tgt_masks)          mdqe/models/matcher.py:182-197 (one clip)
  mask_losses(mask_coeff, proto, tgt, tgt_interinst, n)  mdqe/models/criterion.py:440-473 (one clip, autograd-capable)
  mask_nms_siou(mask_pred)                               mdqe/mdqe.py:394-401
  mask_track_siou(saved_masks, input_masks)              mdqe/tracking/OverTracker.py:92-113
  aligned_bilinear(tensor, factor, sigmoid=False)        mdqe/util/misc.py:485-507 (+ mdqe/mdqe.py:357)
  query_init_sample(encoded_feat, spatial_shapes, level_start_index, coords)
                                                         mdqe/models/transformer_dec.py:170-179 (autograd-capable)
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .ops import _check_inputs, _stream_ptr


def _f32(who, tensors):
    _check_inputs(who, tensors)
    for name, t in tensors:
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: {name} must be float32, got {t.dtype}")


def mask_match_cost(mask_coeff, proto, tgt_masks):
    """cost_bce, cost_dice [Q,G] of one clip: what matcher.py:182-197 computes as
    ``out_mask = einsum('qm,mthw->qthw'); batch_sigmoid_ce_loss(out_mask, tgt); batch_dice_loss(out_mask, tgt)``,
    without materialising out_mask.  mask_coeff [Q,K], proto [K,T,H,W], tgt_masks [G,T,H,W] (same plane as proto)."""
    who = "mask_match_cost"
    tgt_masks = tgt_masks.to(proto.dtype) if tgt_masks.dtype != proto.dtype else tgt_masks        # matcher.py:192 `.to(out_mask)`
    _f32(who, [("mask_coeff", mask_coeff), ("proto", proto), ("tgt_masks", tgt_masks)])
    if mask_coeff.dim() != 2 or proto.shape[0] != mask_coeff.shape[1] or tuple(tgt_masks.shape[1:]) != tuple(proto.shape[1:]):
        raise RuntimeError(f"{who}: expected mask_coeff[Q,K], proto[K,...], tgt_masks[G,...], got {tuple(mask_coeff.shape)} "
                           f"{tuple(proto.shape)} {tuple(tgt_masks.shape)}")
    Q, K = mask_coeff.shape
    G = tgt_masks.shape[0]
    ncols = proto.numel() // max(K, 1)
    lib = _lib.load()
    with torch.cuda.device(proto.device):
        ws = torch.empty(lib.mask_match_cost_workspace_bytes(), dtype=torch.uint8, device=proto.device)
        bce = torch.empty(Q, G, dtype=torch.float32, device=proto.device)
        dice = torch.empty(Q, G, dtype=torch.float32, device=proto.device)
        rc = lib.mask_match_cost(_stream_ptr(proto.device), mask_coeff.data_ptr(), proto.data_ptr(), tgt_masks.data_ptr(), Q, K, G,
                                 ncols, ws.data_ptr(), bce.data_ptr(), dice.data_ptr())
    _lib.check(rc, who)
    return bce, dice


class _MaskLossesFunction(Function):
    """(loss_mask, loss_dice) of up to 32 matched rows of one clip; see mask_losses()."""

    @staticmethod
    def forward(ctx, coeff, proto, tgt, tgt_inter, num_masks):
        who = "mask_losses"
        tensors = [("mask_coeff", coeff), ("proto", proto), ("tgt_masks", tgt)] + ([("tgt_interinst_masks", tgt_inter)] if tgt_inter is not None else [])
        _f32(who, tensors)
        G, K = coeff.shape
        ncols = proto.numel() // max(K, 1)
        if proto.shape[0] != K or tgt.numel() != G * ncols or (tgt_inter is not None and tgt_inter.numel() != G * ncols):
            raise RuntimeError(f"{who}: expected mask_coeff[G,K], proto[K,...], targets[G,...], got {tuple(coeff.shape)} {tuple(proto.shape)} {tuple(tgt.shape)}")
        lib = _lib.load()
        with torch.cuda.device(proto.device):
            ws = torch.empty(lib.mask_losses_workspace_bytes(), dtype=torch.uint8, device=proto.device)
            stats = torch.empty(max(G, 1), 8, dtype=torch.float32, device=proto.device)
            losses = torch.empty(2, dtype=torch.float32, device=proto.device)
            rc = lib.mask_losses_forward(_stream_ptr(proto.device), coeff.data_ptr(), proto.data_ptr(), tgt.data_ptr(),
                                         tgt_inter.data_ptr() if tgt_inter is not None else None, G, K, ncols, float(num_masks),
                                         ws.data_ptr(), stats.data_ptr(), losses.data_ptr())
        _lib.check(rc, who)
        ctx.save_for_backward(coeff, proto, tgt, tgt_inter, stats)
        ctx.num_masks = float(num_masks)
        return losses

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_losses):
        coeff, proto, tgt, tgt_inter, stats = ctx.saved_tensors
        G, K = coeff.shape
        ncols = proto.numel() // max(K, 1)
        grad_losses = grad_losses.contiguous().float()
        lib = _lib.load()
        with torch.cuda.device(proto.device):
            gc = torch.empty_like(coeff)
            gp = torch.empty_like(proto)
            rc = lib.mask_losses_backward(_stream_ptr(proto.device), coeff.data_ptr(), proto.data_ptr(), tgt.data_ptr(),
                                          tgt_inter.data_ptr() if tgt_inter is not None else None, stats.data_ptr(), grad_losses.data_ptr(),
                                          G, K, ncols, ctx.num_masks, gc.data_ptr(), gp.data_ptr())
        _lib.check(rc, "mask_losses (backward)")
        return gc, gp, None, None, None


def mask_losses(mask_coeff, proto, tgt_masks, tgt_interinst_masks, num_masks):
    """loss_mask, loss_dice of SetCriterion.loss_masks (mdqe/models/criterion.py:440-473) for the matched queries of ONE clip:
    mask_coeff [G,K] = outputs["mask_coeff"][b][src_idx], proto [K,T,H,W] = outputs["proto"][b], tgt_masks [G,T,H,W];
    tgt_interinst_masks [G,T,H,W] selects the inter-instance forms (:467-470), None the plain ones (:472-475).
    Autograd-capable (gradients for mask_coeff and proto); the [Q,T,H,W] src_masks of the reference are never materialised.
    Clips of a batch are summed by the caller: every loss is a sum over rows divided by the same num_masks."""
    tgt_masks = tgt_masks.to(proto.dtype).contiguous()
    ti = tgt_interinst_masks.to(proto.dtype).contiguous() if tgt_interinst_masks is not None else None
    mask_coeff, proto = mask_coeff.contiguous(), proto.contiguous()
    G = mask_coeff.shape[0]
    total = None
    for g0 in range(0, max(G, 1), 32):                     # the kernels take up to 32 rows per launch
        part = _MaskLossesFunction.apply(mask_coeff[g0:g0 + 32].contiguous(), proto, tgt_masks[g0:g0 + 32].contiguous(),
                                         ti[g0:g0 + 32].contiguous() if ti is not None else None, num_masks)
        total = part if total is None else total + part
    return total[0], total[1]


def mask_nms_siou(mask_pred):
    """siou [Q,Q] of mdqe/mdqe.py:394-401 from mask_pred [Q,T,H,W] in one pass."""
    who = "mask_nms_siou"
    _f32(who, [("mask_pred", mask_pred)])
    if mask_pred.dim() != 4:
        raise RuntimeError(f"{who}: expected mask_pred[Q,T,H,W], got {tuple(mask_pred.shape)}")
    Q, T, H, W = mask_pred.shape
    lib = _lib.load()
    with torch.cuda.device(mask_pred.device):
        ws = torch.empty(lib.mask_nms_siou_workspace_bytes(), dtype=torch.uint8, device=mask_pred.device)
        siou = torch.empty(Q, Q, dtype=torch.float32, device=mask_pred.device)
        rc = lib.mask_nms_siou(_stream_ptr(mask_pred.device), mask_pred.data_ptr(), Q, T, H, W, ws.data_ptr(), siou.data_ptr())
    _lib.check(rc, who)
    return siou


def mask_track_siou(saved_masks, input_masks):
    """Drop-in for OverTracker._get_siou(saved_masks, input_masks) (mdqe/tracking/OverTracker.py:92-113): [Ns,T,H,W] and
    [Ni,T,H,W] mask probabilities -> hard-mask IoU [Ns,Ni]."""
    who = "mask_track_siou"
    _f32(who, [("saved_masks", saved_masks), ("input_masks", input_masks)])
    if saved_masks.dim() != 4 or input_masks.dim() != 4 or tuple(saved_masks.shape[1:]) != tuple(input_masks.shape[1:]):
        raise RuntimeError(f"{who}: expected [Ns,T,H,W] and [Ni,T,H,W], got {tuple(saved_masks.shape)} {tuple(input_masks.shape)}")
    Ns, T, H, W = saved_masks.shape
    Ni = input_masks.shape[0]
    lib = _lib.load()
    with torch.cuda.device(saved_masks.device):
        ws = torch.empty(lib.mask_nms_siou_workspace_bytes(), dtype=torch.uint8, device=saved_masks.device)
        siou = torch.empty(Ns, Ni, dtype=torch.float32, device=saved_masks.device)
        rc = lib.mask_track_siou(_stream_ptr(saved_masks.device), saved_masks.data_ptr(), input_masks.data_ptr(), Ns, Ni, T, H, W,
                                 ws.data_ptr(), siou.data_ptr())
    _lib.check(rc, who)
    return siou


def aligned_bilinear(tensor, factor, sigmoid=False):
    """Same signature as mdqe/util/misc.py:485 (4-D tensor, integer factor >= 1); ``sigmoid=True`` fuses the ``.sigmoid()`` of
    mdqe/mdqe.py:357 into the same pass."""
    who = "aligned_bilinear"
    assert tensor.dim() == 4
    assert factor >= 1
    assert int(factor) == factor
    factor = int(factor)
    _f32(who, [("tensor", tensor)])
    n, c, h, w = tensor.shape
    lib = _lib.load()
    with torch.cuda.device(tensor.device):
        out = torch.empty(n, c, h * factor, w * factor, dtype=torch.float32, device=tensor.device)
        rc = lib.aligned_bilinear_sigmoid(_stream_ptr(tensor.device), tensor.data_ptr(), n * c, h, w, factor, int(bool(sigmoid)),
                                          out.data_ptr())
    _lib.check(rc, who)
    return out


class _QueryInitSampleFunction(Function):
    @staticmethod
    def forward(ctx, feat, shapes, level_start, coords):
        who = "query_init_sample"
        _f32(who, [("encoded_feat", feat), ("coords", coords)])
        _check_inputs(who, [("spatial_shapes", shapes), ("level_start_index", level_start)])
        if shapes.dtype != torch.int64 or level_start.dtype != torch.int64:
            raise RuntimeError(f"{who}: spatial_shapes and level_start_index must be int64")
        B, S, C = feat.shape
        Q = coords.shape[1]
        if tuple(coords.shape) != (B, Q, 2):
            raise RuntimeError(f"{who}: expected coords[B,Q,2], got {tuple(coords.shape)}")
        lib = _lib.load()
        with torch.cuda.device(feat.device):
            out = torch.empty(B, Q, C, dtype=torch.float32, device=feat.device)
            rc = lib.query_init_sample_forward(_stream_ptr(feat.device), feat.data_ptr(), shapes.data_ptr(), level_start.data_ptr(),
                                               coords.data_ptr(), B, S, C, shapes.shape[0], Q, out.data_ptr())
        _lib.check(rc, who)
        ctx.save_for_backward(feat, shapes, level_start, coords)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        feat, shapes, level_start, coords = ctx.saved_tensors
        B, S, C = feat.shape
        Q = coords.shape[1]
        grad_out = grad_out.contiguous()
        lib = _lib.load()
        with torch.cuda.device(feat.device):
            gf = torch.empty_like(feat)
            gc = torch.empty_like(coords)
            rc = lib.query_init_sample_backward(_stream_ptr(feat.device), feat.data_ptr(), shapes.data_ptr(), level_start.data_ptr(),
                                                coords.data_ptr(), grad_out.data_ptr(), B, S, C, shapes.shape[0], Q, gf.data_ptr(),
                                                gc.data_ptr())
        _lib.check(rc, "query_init_sample (backward)")
        return gf, None, None, gc


def query_init_sample(encoded_feat, spatial_shapes, level_start_index, coords):
    """mean over the pyramid levels of ``F.grid_sample(feat_l, 2*coords-1, bilinear, padding_mode='border',
    align_corners=False)`` (transformer_dec.py:170-179): encoded_feat [B,S,C], coords [B,Q,2] in (x, y) -> [B,Q,C]."""
    return _QueryInitSampleFunction.apply(encoded_feat.contiguous(), spatial_shapes.contiguous(), level_start_index.contiguous(),
                                          coords.contiguous())
