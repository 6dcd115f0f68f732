"""Shared test helpers: an autograd Function backed by the plain-C oracle (CPU) so that host-side
module logic can be checked without a GPU, plus error metrics and golden loading."""
import ast
import os

import numpy as np
import torch

from oracle import msda_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def nerr(a, b):
    """normalised max error  ||a-b||_inf / ||b||_inf  (SURVEY 8d tolerance metric)."""
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max() if b.size else 0.0
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0)) if a.size else 0.0


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def ctor_kwargs(z):
    return dict(ast.literal_eval(str(z["ctor"])))


class OracleMSDAFunction(torch.autograd.Function):
    """CPU stand-in with the MSDeformAttnFunction signature, computed by oracle/msda_oracle.c."""

    @staticmethod
    def forward(ctx, value, shapes, level_start, loc, aw, im2col_step):
        ctx.save_for_backward(value, shapes, level_start, loc, aw)
        out = O.msda_forward(value.detach().numpy(), shapes.numpy(), loc.detach().numpy(), aw.detach().numpy(),
                             level_start.numpy())
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_out):
        value, shapes, level_start, loc, aw = ctx.saved_tensors
        gv, gl, ga = O.msda_backward(value.detach().numpy(), shapes.numpy(), loc.detach().numpy(), aw.detach().numpy(),
                                     grad_out.contiguous().numpy(), level_start.numpy())
        return torch.from_numpy(gv), None, None, torch.from_numpy(gl), torch.from_numpy(ga), None


def module_from_golden(z, module_cls):
    mod = module_cls(**ctor_kwargs(z))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    mod.load_state_dict(sd, strict=True)
    return mod


def module_inputs(z, device="cpu"):
    mask = torch.from_numpy(z["padding_mask"])
    mask = mask.to(device) if mask.numel() else None
    return (torch.from_numpy(z["query"]).to(device).requires_grad_(True),
            torch.from_numpy(z["reference_points"]).to(device),
            torch.from_numpy(z["input_flatten"]).to(device).requires_grad_(True),
            torch.from_numpy(z["spatial_shapes"]).to(device), mask)
