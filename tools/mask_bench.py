#!/usr/bin/env python
"""Mask contraction only: ours (fp32 3xTF32 / bf16 / fp16 tcgen05 kernels) vs torch.einsum (cuBLAS) at the BASELINE configs[4]
sweep shapes, forward and backward.  CUDA events around each call, L2 flushed between launches; median of --iters."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mdqe_cvpr2023_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 25
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = []
for Q, T, plane, K in ((196, 4, (96, 160), 32), (196, 4, (160, 288), 32), (300, 8, (96, 160), 32), (196, 3, (96, 160), 24), (100, 2, (96, 160), 32)):
    n = T * plane[0] * plane[1]
    coeff = torch.tanh(torch.randn(1, Q, K, device="cuda"))
    proto = torch.randn(1, K, T, *plane, device="cuda")
    go = torch.randn(1, Q, T, *plane, device="cuda")
    row = {"shape": f"Q{Q}_K{K}_N{n}"}
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16), ("f16", torch.float16)):
        c, p = coeff.to(dt), proto.to(dt)
        row[f"ours_{name}"] = timed(lambda: ops.mask_logits_forward(c, p))
        row[f"einsum_{name}"] = timed(lambda: torch.einsum("bqm,bmthw->bqthw", c, p))
    row["ours_bwd_f32"] = timed(lambda: ops.mask_logits_backward(coeff, proto, go))
    res.append(row)
    print("  ".join(f"{k}: {v:6.1f}" if isinstance(v, float) else f"{v:18s}" for k, v in row.items()), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mask_bench.json"), "w"), indent=1)
