"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE'S OWN PYTHON.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed, unmodified, from /root/reference:
  * mdqe/models/ops/functions/ms_deform_attn_func.py:45-65  ms_deform_attn_core_pytorch
    (+ torch autograd through it for the three gradients)
  * mdqe/models/ops/modules/ms_deform_attn.py  MSDeformAttn (spatial / temporal,
    pred_offsets True / False) with MSDeformAttnFunction pointed at that pure-torch function
  * the literal mask einsum 'bqm,bmthw->bqthw' (mdqe/models/matcher.py:182)

The compiled module `MultiScaleDeformableAttention` the reference imports at
ms_deform_attn_func.py:19 is stubbed in sys.modules (it is never called), and `mdqe`,
`mdqe.models`, `mdqe.models.ops` are pre-registered as bare namespace packages so that
mdqe/__init__.py (detectron2) never runs.

Outputs: one .npz per case (inputs + reference outputs), all small enough to commit.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def import_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    for name, sub in (("mdqe", "mdqe"), ("mdqe.models", "mdqe/models"), ("mdqe.models.ops", "mdqe/models/ops")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = mod
    import mdqe.models.ops.functions.ms_deform_attn_func as func
    import mdqe.models.ops.modules.ms_deform_attn as module
    return func, module


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def run_core(func, value, shapes, loc, aw, grad_out):
    value = value.clone().requires_grad_(True)
    loc = loc.clone().requires_grad_(True)
    aw = aw.clone().requires_grad_(True)
    out = func.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
    out.backward(grad_out)
    return out.detach(), value.grad, loc.grad, aw.grad


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def core_case(func, name, value, shapes, loc, aw, seed_go, extra=None):
    g = torch.Generator().manual_seed(seed_go)
    N, Lq, M = loc.shape[0], loc.shape[1], loc.shape[2]
    D = value.shape[3]
    grad_out = torch.rand(N, Lq, M * D, generator=g, dtype=value.dtype) - 0.5
    out, gv, gl, ga = run_core(func, value, shapes, loc, aw, grad_out)
    save(name, value=value, shapes=shapes, level_start=lsi_of(shapes), loc=loc, aw=aw, grad_out=grad_out,
         out=out, grad_value=gv, grad_loc=gl, grad_aw=ga, **(extra or {}))


def main():
    func, module = import_reference()

    # ---- 1. the reference's only test fixture, ops/test.py:21-36 (seed 3), double and float, D sweep
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    for D in (2, 4, 30, 32, 64, 71):
        torch.manual_seed(3)
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
        core_case(func, f"ref_fixture_D{D}_f64", value.double(), shapes, loc.double(), aw.double(), 100 + D)
        if D in (2, 32):
            core_case(func, f"ref_fixture_D{D}_f32", value, shapes, loc, aw, 100 + D)

    # ---- 2. MDQE-like pyramids (R50: D=32, Swin-L: D=24), samples partly outside [0,1]
    for name, D, dt in (("pyramid_D32_f32", 32, torch.float32), ("pyramid_D32_f64", 32, torch.float64),
                        ("pyramid_D24_f32", 24, torch.float32)):
        g = torch.Generator().manual_seed(7)
        shapes = torch.as_tensor([(8, 10), (4, 5), (2, 3), (1, 2)], dtype=torch.long)
        S = int(shapes.prod(1).sum())
        N, M, Lq, L, P = 2, 8, 23, 4, 4
        value = torch.randn(N, S, M, D, generator=g, dtype=dt)
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=dt) * 1.4 - 0.2
        aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, dtype=dt), -1).view(N, Lq, M, L, P)
        core_case(func, name, value, shapes, loc, aw, 11)

    # ---- 3. temporal mode: "levels" are T frames of one pyramid level (ms_deform_attn.py:219-233)
    g = torch.Generator().manual_seed(9)
    T, H, W = 3, 4, 6
    shapes = torch.as_tensor([(H, W)] * T, dtype=torch.long)
    N, M, D, Lq, P = 2, 8, 32, 11, 4
    value = torch.randn(N, T * H * W, M, D, generator=g)
    loc = torch.rand(N, Lq, M, T, P, 2, generator=g) * 1.2 - 0.1
    aw = torch.softmax(torch.randn(N, Lq, M, T * P, generator=g), -1).view(N, Lq, M, T, P)
    core_case(func, "temporal_T3_f32", value, shapes, loc, aw, 13)

    # ---- 4. edge coordinates: pixel centres, exact borders, just outside, far outside
    shapes = torch.as_tensor([(4, 5)], dtype=torch.long)
    pts = []
    for x in (-0.3, -0.1, 0.0, 0.1, 0.5, 0.9, 1.0, 1.1, 1.3, 3.0):
        for y in (-0.125, 0.0, 0.125, 0.375, 0.875, 1.0, 1.125):
            pts.append((x, y))
    loc = torch.tensor(pts, dtype=torch.float64).view(1, len(pts), 1, 1, 1, 2)
    g = torch.Generator().manual_seed(21)
    value = torch.randn(1, 20, 1, 8, generator=g, dtype=torch.float64)
    aw = torch.rand(1, len(pts), 1, 1, 1, generator=g, dtype=torch.float64) + 0.5
    core_case(func, "edge_coords_f64", value, shapes, loc, aw, 17)

    # ---- 5. module level: the unmodified reference MSDeformAttn on the pure-torch function
    class _OracleFn:
        @staticmethod
        def apply(value, shapes, lsi, loc, aw, step):
            return func.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
    module.MSDeformAttnFunction = _OracleFn

    def module_case(name, ctor_kw, make_inputs):
        torch.manual_seed(5)
        mod = module.MSDeformAttn(**ctor_kw)
        # move the weights off their (mostly zero) initial values so every branch matters
        g = torch.Generator().manual_seed(31)
        with torch.no_grad():
            for p in mod.parameters():
                p.add_(0.2 * torch.randn(p.shape, generator=g))
        query, ref, inp, shapes, mask = make_inputs(g)
        query = query.requires_grad_(True)
        inp = inp.requires_grad_(True)
        out = mod(query, ref, inp, shapes, mask)
        go = torch.rand(out.shape, generator=g) - 0.5
        out.backward(go)
        sd = {"sd." + k: v for k, v in mod.state_dict().items()}
        grads = {"gp." + k: p.grad for k, p in mod.named_parameters()}
        save(name, query=query, reference_points=ref, input_flatten=inp, spatial_shapes=shapes,
             padding_mask=mask if mask is not None else np.zeros(0, dtype=bool),
             out=out, grad_out=go, grad_query=query.grad, grad_input=inp.grad,
             ctor=np.array(repr(sorted(ctor_kw.items()))), **sd, **grads)

    shapes4 = torch.as_tensor([(8, 12), (4, 6), (2, 3), (1, 2)], dtype=torch.long)
    S4 = int(shapes4.prod(1).sum())

    def enc_inputs(g):
        B = 2
        query = torch.randn(B, S4, 64, generator=g)
        ref = torch.cat([torch.rand(B, S4, 2, generator=g), torch.full((B, S4, 2), 0.1)], -1)
        inp = torch.randn(B, S4, 64, generator=g)
        mask = torch.rand(B, S4, generator=g) < 0.1
        return query, ref, inp, shapes4, mask

    def dec_inputs(g):
        B, Q = 2, 9
        query = torch.randn(B, Q, 64, generator=g)
        ref = torch.cat([torch.rand(B, Q, 2, generator=g), torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05], -1)
        inp = torch.randn(B, S4, 64, generator=g)
        return query, ref, inp, shapes4, None

    def temporal_inputs(g):
        B, Q, T = 2, 9, 3
        query = torch.randn(B, Q, 64, generator=g)
        ref = torch.cat([torch.rand(B, Q, 2, generator=g), torch.rand(B, Q, 2, generator=g) * 0.3 + 0.05], -1)
        inp = torch.randn(B, T, S4, 64, generator=g)
        mask = torch.rand(B, T, S4, generator=g) < 0.1
        return query, ref, inp, shapes4, mask

    module_case("module_spatial_pred", dict(d_model=64, n_levels=4, n_heads=4, n_points=4, pred_offsets=True,
                                            mode="spatial"), enc_inputs)
    module_case("module_spatial_grid", dict(d_model=64, n_levels=4, n_heads=4, n_points=4, pred_offsets=False,
                                            mode="spatial"), dec_inputs)
    module_case("module_temporal_grid", dict(d_model=64, n_levels=4, n_heads=4, n_points=4, n_frames=3,
                                             pred_offsets=False, mode="temporal"), temporal_inputs)

    # ---- 6. mask contraction: the literal einsum at matcher.py:182 / criterion.py:440 / mdqe.py:384
    g = torch.Generator().manual_seed(41)
    for name, B, Q, K, T, H, W in (("mask_einsum_K32", 2, 21, 32, 2, 6, 10), ("mask_einsum_K24", 1, 13, 24, 3, 5, 7)):
        coeff = torch.tanh(torch.randn(B, Q, K, generator=g)).requires_grad_(True)
        proto = torch.randn(B, K, T, H, W, generator=g).requires_grad_(True)
        out = torch.einsum('bqm,bmthw->bqthw', coeff, proto)
        go = torch.rand(out.shape, generator=g) - 0.5
        out.backward(go)
        save(name, coeff=coeff, proto=proto, out=out, grad_out=go, grad_coeff=coeff.grad, grad_proto=proto.grad)


if __name__ == "__main__":
    main()
