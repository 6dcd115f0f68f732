#!/usr/bin/env python
"""Timing-only "what if" experiment on the encoder shape: how fast would the fast2 kernels be if the gathers /
grad_value reductions of some pyramid levels were handled elsewhere (results are WRONG by construction)?
Builds a separate library with -DMSDA_DBG_MASK (never the product library) and loads it through MSDA_B200_LIB.

    python tools/whatif_bench.py --build     # here (no GPU needed)
    python tools/whatif_bench.py             # on the GPU box
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "mdqe_cvpr2023_b200", "libmsda_b200_dbg.so")
sys.path.insert(0, ROOT)

if "--build" in sys.argv:
    from mdqe_cvpr2023_b200 import build as b
    cmd = ["/usr/local/cuda/bin/nvcc"] + b.NVCC_FLAGS + ["-DMSDA_DBG_MASK", "-o", DBG] + b.SOURCES
    subprocess.check_call(cmd, cwd=b.CSRC)
    print(DBG)
    sys.exit(0)

os.environ["MSDA_B200_LIB"] = DBG
import ctypes  # noqa: E402

import torch  # noqa: E402

from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, make_inputs, to_cuda  # noqa: E402

lib = _lib.load()
lib.msda_debug_set_mask.argtypes = [ctypes.c_int]
flush = torch.empty(128 * 1024 * 1024, device="cuda")


def timed(fn, iters=15):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for dist in ("local", "uniform"):
    inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, dist=dist, seed=0))
    a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
    for name, mask in [("full", 0), ("fwd_no_l3", 0x8), ("fwd_no_l23", 0xC), ("fwd_no_l123", 0xE), ("fwd_none", 0xF)]:
        lib.msda_debug_set_mask(mask)
        res[f"{dist}/fwd/{name}"] = timed(lambda: ops.ms_deform_attn_forward(*a, 64))
    for name, mask in [("full", 0), ("nored_l3", 0x80), ("nored_l23", 0xC0), ("nored_l123", 0xE0), ("nored_all", 0xF0),
                       ("nogather_nored_l3", 0x800), ("nogather_nored_l23", 0xC00), ("nothing", 0xF00), ("nored_l0", 0x10), ("nored_l01", 0x30)]:
        lib.msda_debug_set_mask(mask)
        res[f"{dist}/bwd/{name}"] = timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64))
    lib.msda_debug_set_mask(0)
for k, v in res.items():
    print(f"{k:40s} {v:8.1f} us")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "whatif_bench.json"), "w"), indent=1)
