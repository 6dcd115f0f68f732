#!/usr/bin/env python
"""Tiny launch target for `ncu`: runs ONE kind of call a few times so that -k/-s/-c select it easily.

    python tools/profile_target.py enc_fwd|enc_bwd|dec_fwd|dec_bwd|mask_fwd|mask_bwd|lin_fwd|lin_dgrad|lin_wgrad [--dist local] [--dtype fp32] [--reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mdqe_cvpr2023_b200 import ops  # noqa: E402
from tests.gpu_util import R50_360, make_inputs, to_cuda  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--dist", default="local")
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--opt", default="", help="library options, e.g. pair_map=2,chunk_pairs=64")
args = ap.parse_args()
if args.opt:
    from mdqe_cvpr2023_b200 import _lib
    for kv in args.opt.split(","):
        k, v = kv.split("=")
        _lib.set_option(k, int(v))

if args.what.startswith("lin"):
    x = torch.randn(4 * 5100, 256, device="cuda")
    w = torch.randn(256, 256, device="cuda") / 16
    b = torch.randn(256, device="cuda")
    gy = torch.randn(4 * 5100, 256, device="cuda")
    for _ in range(args.reps):
        if args.what == "lin_fwd":
            ops.tc_linear_forward(x, w, b)
        elif args.what == "lin_dgrad":
            ops.tc_linear_backward(gy, x, w, True, False)
        else:
            ops.tc_linear_backward(gy, x, w, False, True, True)
elif args.what.startswith("mask"):
    coeff = torch.tanh(torch.randn(1, 196, 32, device="cuda"))
    proto = torch.randn(1, 32, 4, 96, 160, device="cuda")
    go = torch.randn(1, 196, 4, 96, 160, device="cuda")
    if args.dtype == "bf16":
        coeff, proto = coeff.bfloat16(), proto.bfloat16()
    for _ in range(args.reps):
        if args.what == "mask_fwd":
            ops.mask_logits_forward(coeff, proto)
        else:
            ops.mask_logits_backward(coeff, proto, go)
else:
    enc = args.what.startswith("enc")
    inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, Lq=None if enc else 196, dist=args.dist, seed=0))
    if args.dtype == "bf16":
        inp = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in inp.items()}
    a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
    for _ in range(args.reps):
        if args.what.endswith("fwd"):
            ops.ms_deform_attn_forward(*a, 64)
        else:
            ops.ms_deform_attn_backward(*a, inp["grad_out"], 64)
torch.cuda.synchronize()
