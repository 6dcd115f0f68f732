#!/usr/bin/env python
"""Encoder-shape backward: write-combining knobs (wc_max_cells, wc_chunk) x sampling distribution; CUDA events, L2 flushed by a read."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, make_inputs, oracle_all, to_cuda  # noqa: E402

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
shapes = {"R50_360": (R50_360, 32), "R50_720": ([(80, 144), (40, 72), (20, 36), (10, 18)], 32)}
for sname, (pyr, D) in shapes.items():
    for dist in ("local", "uniform"):
        inp = to_cuda(make_inputs(4, pyr, 8, D, 4, dist=dist, seed=0))
        a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"], inp["grad_out"])
        ref = None
        for cells, chunk in [(0, 512), (1, 512), (64, 512), (256, 150), (256, 512), (1024, 512)]:
            if sname != "R50_360" and chunk != 512:
                continue
            _lib.set_option("wc_max_cells", cells)
            _lib.set_option("wc_chunk", chunk)
            t = timed(lambda: ops.ms_deform_attn_backward(*a, 64))
            gv, gl, ga = ops.ms_deform_attn_backward(*a, 64)
            if ref is None:
                ref = gv.clone()
            err = float((gv - ref).abs().max() / ref.abs().max())
            res[f"{sname}/{dist}/cells{cells}/chunk{chunk}"] = {"us": t, "gv_vs_uncached": err}
            print(f"{sname:8s} {dist:8s} wc_max_cells={cells:5d} wc_chunk={chunk:5d}  {t:8.1f} us   grad_value vs uncached {err:.2e}")
_lib.set_option("wc_max_cells", 256)
_lib.set_option("wc_chunk", 512)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "wc_sweep.json"), "w"), indent=1)
