#!/usr/bin/env python
"""A/B of the backward's in-level reduction merging (option bwd_merge): encoder shapes, CUDA events, L2 flushed by a read.
Also checks that both settings give the same gradients (max normalised difference)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, R50_720, make_inputs, to_cuda  # noqa: E402

flush = torch.ones(160 * 1024 * 1024, device="cuda")


def timed(fn, iters=12):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def nerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


modes = [int(x) for x in (sys.argv[1:] or ["0", "1"])]
res = {}
for sname, pyr, D, N in (("R50_360", R50_360, 32, 4), ("R50_720", R50_720, 32, 4), ("swinl_360", R50_360, 24, 3)):
    for dist in ("local", "uniform"):
        for dt in (torch.float32, torch.bfloat16):
            inp = to_cuda(make_inputs(N, pyr, 8, D, 4, dist=dist, seed=0, dtype=dt))
            a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
            out, base = {}, None
            for mode in modes:
                _lib.set_option("bwd_merge", mode)
                out[mode] = timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64))
                g = ops.ms_deform_attn_backward(*a, inp["grad_out"], 64)
                if base is None:
                    base = g
                else:
                    out[f"nerr{mode}"] = max(nerr(x, y) for x, y in zip(g, base))
            _lib.set_option("bwd_merge", 1)
            res[f"{sname}/{dist}/{str(dt)[6:]}"] = out
            print(f"{sname:10s} {dist:8s} {str(dt)[6:]:9s}", "  ".join((f"{k}: {v:.3g}" if "nerr" in k else f"{k}: {v:6.1f}") if isinstance(k, str) else f"bwd{k}: {v:6.1f}" for k, v in out.items()))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({str(k): v for k, v in res.items()}, open(os.path.join(ROOT, "gpurun_out", "merge_ab.json"), "w"), indent=1, default=str)
