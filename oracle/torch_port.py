"""PyTorch (CPU) port of the reference's own CPU path -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference has no native CPU kernel (ops/src/cpu/ms_deform_attn_cpu.cpp:17-40 only raise); its CPU
implementation of the hot path is the pure-PyTorch function ms_deform_attn_core_pytorch
(/root/reference/mdqe/models/ops/functions/ms_deform_attn_func.py:45-65: F.grid_sample per level, then a
weighted sum) with gradients from autograd, and the literal einsum for the mask contraction
(mdqe/models/matcher.py:182).  /root/reference does not exist on the GPU box, so bench.py's
`--impl reference` arm and `cpu_baseline` leg time THIS restatement there (cpu_baseline.kind = "port").
It uses all host threads torch is given.  tests/test_oracle_golden.py pins it against fixtures recorded
from the reference's own function.

Generalisation over the reference function: `level_start` may be passed explicitly so that "levels" can
be arbitrary windows of the value rows (temporal mode of mdqe_cvpr2023_b200.modules).
"""
import torch
import torch.nn.functional as F


def msda_core_torch(value, spatial_shapes, sampling_locations, attention_weights, level_start=None):
    """value [N,S,M,D], spatial_shapes [L,2], sampling_locations [N,Lq,M,L,P,2] in [0,1],
    attention_weights [N,Lq,M,L,P]  ->  [N, Lq, M*D]."""
    n, _, heads, dim = value.shape
    _, n_query, _, n_levels, n_points, _ = sampling_locations.shape
    hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    if level_start is None:
        starts, acc = [], 0
        for h, w in hw:
            starts.append(acc)
            acc += h * w
    else:
        starts = [int(s) for s in level_start.tolist()]
    # grid_sample wants [-1, 1] coordinates and channel-first images, one image per (batch, head)
    grid = sampling_locations * 2 - 1
    per_level = []
    for lvl, (h, w) in enumerate(hw):
        rows = value[:, starts[lvl]:starts[lvl] + h * w]                       # N, h*w, M, D
        image = rows.permute(0, 2, 3, 1).reshape(n * heads, dim, h, w)
        grid_l = grid[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * heads, n_query, n_points, 2)   # contiguous copy
        per_level.append(F.grid_sample(image, grid_l, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))                    # N*M, D, Lq, P
    sampled = torch.stack(per_level, dim=-2).flatten(-2)                        # N*M, D, Lq, L*P
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * heads, 1, n_query, n_levels * n_points)
    out = (sampled * weights).sum(-1)                                           # N*M, D, Lq
    return out.view(n, heads * dim, n_query).transpose(1, 2).contiguous()


def msda_fwd_bwd_torch(value, spatial_shapes, sampling_locations, attention_weights, grad_out, level_start=None):
    """forward + autograd backward; returns (out, grad_value, grad_loc, grad_aw)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    aw = attention_weights.detach().clone().requires_grad_(True)
    out = msda_core_torch(v, spatial_shapes, loc, aw, level_start)
    out.backward(grad_out.view_as(out))
    return out.detach(), v.grad, loc.grad, aw.grad


def mask_logits_torch(coeff, proto):
    return torch.einsum("bqm,bmthw->bqthw", coeff, proto)
