#!/usr/bin/env python
"""Module-level timing: MSDeformAttn forward+backward at the R50_ovis_360 shapes:
   fused_tc : this package's module: softmax/location arithmetic inside the kernel, grouped temporal launch, Linear layers as
              3xTF32 tensor-core GEMMs (the default configuration)
   fused    : the same with the Linear layers on torch (cuBLAS fp32, TF32 off like the reference)
   unfused  : this package's module running the reference's op sequence on our kernels
   refcuda  : the same op sequence on the reference's own CUDA extension (oracle/_ref), per-level loop + .contiguous()
each eager (host dispatch included) and, `*_graph_us`, as a replayed CUDA graph of the whole forward+backward (GPU time only).
Writes gpurun_out/module_bench.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mdqe_cvpr2023_b200.modules as M  # noqa: E402
from mdqe_cvpr2023_b200 import functions, ops  # noqa: E402

PYR = [(48, 80), (24, 40), (12, 20), (6, 10)]
S = sum(h * w for h, w in PYR)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def timed_graph(step, iters=20):
    """capture forward+backward once, time replays"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    return timed(g.replay, iters)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_ext = None
    try:
        from oracle import build_ref_cuda
        ref_ext = build_ref_cuda.load()
    except Exception as e:  # noqa: BLE001
        print("reference CUDA extension unavailable:", e)
    shapes = torch.tensor(PYR, device="cuda")
    res = {}
    torch.manual_seed(0)
    cases = {
        "encoder_self_attn": (M.MSDeformAttn(256, 4, 8, 4, pred_offsets=True, mode="spatial").cuda(),
                              torch.randn(4, S, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")),
        "decoder_frame_attn": (M.MSDeformAttn(256, 4, 8, 4, pred_offsets=False, mode="spatial").cuda(),
                               torch.randn(4, 196, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")),
        "decoder_clip_attn": (M.MSDeformAttn(256, 4, 8, 4, n_frames=4, pred_offsets=False, mode="temporal").cuda(),
                              torch.randn(1, 196, 256, device="cuda"), torch.randn(1, 4, S, 256, device="cuda")),
    }
    for name, (mod, q, x) in cases.items():
        Q = q.shape[1]
        ref = torch.cat([torch.rand(q.shape[0], Q, 2, device="cuda"), torch.full((q.shape[0], Q, 2), 0.1, device="cuda")], -1)
        q.requires_grad_(True)
        x.requires_grad_(True)

        def step():
            mod.zero_grad(set_to_none=True)
            q.grad = x.grad = None
            mod(q, ref, x, shapes, None).sum().backward()
        def step_graph():                              # same work without python-side grad resets (static graph memory)
            return torch.autograd.grad(mod(q, ref, x, shapes, None).sum(), (q, x) + tuple(mod.parameters()))
        row = {}
        mod.fused_prologue, mod.tc_linear = True, True
        row["fused_tc_us"] = timed(step)
        row["fused_tc_graph_us"] = timed_graph(step_graph)
        mod.tc_linear = False
        row["fused_us"] = timed(step)
        row["fused_graph_us"] = timed_graph(step_graph)
        mod.fused_prologue = False
        row["unfused_us"] = timed(step)
        row["unfused_graph_us"] = timed_graph(step_graph)
        if ref_ext is not None:
            # reference op sequence on the reference kernels: plain Function per level with .contiguous() copies
            class RefFn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, value, sh, ls, loc, aw, step_):
                    ctx.save_for_backward(value, sh, ls, loc, aw)
                    return ref_ext.ms_deform_attn_forward(value, sh, ls, loc, aw, 64)

                @staticmethod
                def backward(ctx, go):
                    value, sh, ls, loc, aw = ctx.saved_tensors
                    gv, gl, ga = ref_ext.ms_deform_attn_backward(value, sh, ls, loc, aw, go.contiguous(), 64)
                    return gv, None, None, gl, ga, None
            orig_fn, orig_grouped = M.MSDeformAttnFunction, ops.grouped_supported
            M.MSDeformAttnFunction = RefFn
            M.ops.grouped_supported = lambda *a: False
            try:
                row["refcuda_us"] = timed(step)
            finally:
                M.MSDeformAttnFunction = orig_fn
                M.ops.grouped_supported = orig_grouped
        mod.fused_prologue, mod.tc_linear = True, True
        res[name] = row
        print(name, {k: round(v, 1) for k, v in row.items()}, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "module_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
