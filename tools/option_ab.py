#!/usr/bin/env python
"""A/B of library tuning options on the encoder-shape sampling kernels (forward and backward), CUDA events, L2 flushed between
launches.  Every setting is also compared with the first one (max normalised difference of all outputs / gradients).

    python tools/option_ab.py "pair_map=1" "pair_map=2" "pair_map=2,chunk_pairs=64" [--shapes R50_360,R50_720,swinl_360]
                              [--dists local,uniform] [--dtypes float32,bfloat16] [--tag NAME]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, R50_720, make_inputs, to_cuda  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("settings", nargs="+")
ap.add_argument("--shapes", default="R50_360")
ap.add_argument("--dists", default="local,uniform")
ap.add_argument("--dtypes", default="float32")
ap.add_argument("--tag", default="option_ab")
ap.add_argument("--iters", type=int, default=15)
args = ap.parse_args()

flush = torch.ones(160 * 1024 * 1024, device="cuda")
SHAPES = {"R50_360": (R50_360, 32, 4), "R50_720": (R50_720, 32, 4), "swinl_360": (R50_360, 24, 3)}


def timed(fn, iters):
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


def apply(setting, touched):
    for kv in setting.split(","):
        if not kv or kv == "default":
            continue
        k, v = kv.split("=")
        touched.setdefault(k, _lib.get_option(k) if hasattr(_lib, "get_option") else 0)
        _lib.set_option(k, int(v))


res = {}
for sname in args.shapes.split(","):
    pyr, D, N = SHAPES[sname]
    for dist in args.dists.split(","):
        for dtn in args.dtypes.split(","):
            dt = getattr(torch, dtn)
            inp = to_cuda(make_inputs(N, pyr, 8, D, 4, dist=dist, seed=0, dtype=dt))
            a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
            base = None
            for setting in args.settings:
                touched = {}
                apply(setting, touched)
                f = timed(lambda: ops.ms_deform_attn_forward(*a, 64), args.iters)
                b = timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), args.iters)
                outs = (ops.ms_deform_attn_forward(*a, 64),) + tuple(ops.ms_deform_attn_backward(*a, inp["grad_out"], 64))
                torch.cuda.synchronize()
                for k, v in touched.items():
                    _lib.set_option(k, v)
                if base is None:
                    base, err = outs, 0.0
                else:
                    err = max(nerr(x, y) for x, y in zip(outs, base))
                res[f"{sname}/{dist}/{dtn}/{setting}"] = dict(fwd_us=f, bwd_us=b, nerr_vs_first=err)
                print(f"{sname:10s} {dist:8s} {dtn:9s} {setting:36s} fwd {f:7.1f}  bwd {b:7.1f}  nerr {err:.2g}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", args.tag + ".json"), "w"), indent=1)
