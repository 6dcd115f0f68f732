#!/usr/bin/env python
"""Encoder-shape forward / backward timing (fp32 + bf16, local + uniform), CUDA events, L2 flushed by a READ of a 640 MB buffer."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import ops  # noqa: E402
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


res = {}
for sname, pyr, D in (("R50_360", R50_360, 32), ("R50_720", R50_720, 32), ("swinl_360", R50_360, 24)):
    for dist in ("local", "uniform"):
        for dt in (torch.float32, torch.bfloat16):
            inp = to_cuda(make_inputs(4 if sname != "swinl_360" else 3, pyr, 8, D, 4, dist=dist, seed=0, dtype=dt))
            a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
            f = timed(lambda: ops.ms_deform_attn_forward(*a, 64))
            b = timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64))
            res[f"{sname}/{dist}/{str(dt)[6:]}"] = {"fwd_us": f, "bwd_us": b}
            print(f"{sname:10s} {dist:8s} {str(dt)[6:]:9s} fwd {f:7.1f} us   bwd (incl. memset) {b:7.1f} us")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bwd_quick.json"), "w"), indent=1)
