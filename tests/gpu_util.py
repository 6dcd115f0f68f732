"""Helpers for the -m gpu parity tests: seeded synthetic inputs and oracle comparison."""
import numpy as np
import torch

from oracle import msda_oracle as O

R50_360 = [(48, 80), (24, 40), (12, 20), (6, 10)]      # S = 5100
R50_720 = [(80, 144), (40, 72), (20, 36), (10, 18)]    # S = 15300


def level_start(shapes):
    sizes = shapes[:, 0] * shapes[:, 1]
    return torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])


def pixel_reference_points(shapes_list):
    """centres of every pyramid cell, normalised: the encoder's reference points (models/misc.py:19-28)."""
    pts = []
    for H, W in shapes_list:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5,
                                indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
    return torch.cat(pts)


def make_inputs(N, shapes_list, M, D, P, Lq=None, dist="uniform", seed=0, dtype=torch.float32):
    """value ~ randn; loc uniform [0,1) | local (ref + 0.05 randn, clamped to [-0.1, 1.1]) | wide
    ([-0.3, 1.3], many samples outside); aw = softmax(randn); grad_out ~ randn.  CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(shapes_list, dtype=torch.long)
    L = len(shapes_list)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    Lq = S if Lq is None else Lq
    value = torch.randn(N, S, M, D, generator=g)
    if dist == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    elif dist == "wide":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.6 - 0.3
    elif dist == "local":
        ref = pixel_reference_points(shapes_list)
        if Lq != S:
            ref = ref[torch.randint(0, S, (Lq,), generator=g)]
        loc = ref.view(1, Lq, 1, 1, 1, 2) + 0.05 * torch.randn(N, Lq, M, L, P, 2, generator=g)
        loc = loc.clamp(-0.1, 1.1)
    else:
        raise ValueError(dist)
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value.to(dtype), shapes=shapes, level_start=level_start(shapes), loc=loc.to(dtype),
                aw=aw.to(dtype), grad_out=grad_out.to(dtype))


def oracle_all(inp):
    """fp32/fp64 oracle forward + backward on whatever precision the inputs carry (bf16 is widened)."""
    def np_(t):
        t = t.detach().cpu()
        return (t.float() if t.dtype == torch.bfloat16 else t).numpy()
    v, loc, aw, go = np_(inp["value"]), np_(inp["loc"]), np_(inp["aw"]), np_(inp["grad_out"])
    sh, ls = inp["shapes"].cpu().numpy(), inp["level_start"].cpu().numpy()
    out = O.msda_forward(v, sh, loc, aw, ls)
    gv, gl, ga = O.msda_backward(v, sh, loc, aw, go, ls)
    return out, gv, gl, ga


def to_cuda(inp):
    return {k: v.cuda() for k, v in inp.items()}


def kink_mask(inp, eps=1e-6):
    """[N,Lq,M,L,P] bool: samples whose pixel coordinate sits (numerically) on a pixel centre line.
    grad_sampling_loc is discontinuous there and the reference's own two code paths disagree: grid_sample
    unnormalises ((g+1)*W-1)/2 while the CUDA kernel computes loc*W-0.5 (ms_deform_im2col_cuda.cuh:285-286), so the
    last bit decides which cell's slope is returned.  Such samples are excluded from grad_loc comparisons."""
    loc = inp["loc"].detach().float().cpu()
    shapes = inp["shapes"].cpu()
    W = shapes[:, 1].view(1, 1, 1, -1, 1).float()
    H = shapes[:, 0].view(1, 1, 1, -1, 1).float()
    px, py = loc[..., 0] * W - 0.5, loc[..., 1] * H - 0.5
    near = lambda t, size: (t - t.round()).abs() < eps * size
    inside = (px >= -1 - 1e-4) & (px <= W + 1e-4) & (py >= -1 - 1e-4) & (py <= H + 1e-4)   # the borders -1 and W/H are kinks too
    return ((near(px, W) | near(py, H)) & inside).numpy()
