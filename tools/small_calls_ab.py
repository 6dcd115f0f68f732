#!/usr/bin/env python
"""Decoder-sized calls (196 queries): forward schedule A/B (fwd_variant 0 = register-lean, 3 = batched gathers) and chunk sizes.
CUDA-graph replay of 20 back-to-back launches on separate buffers (the calls are launch/latency bound, so events around one launch
mostly time the launch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, make_inputs, to_cuda  # noqa: E402

REP = 20


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / REP)
    return sorted(ts)[len(ts) // 2]


# 8 distinct input sets (8 x 21 MB of value > the 126 MB L2 together with loc/aw/out): every launch finds its rows in HBM
sets = []
for k in range(8):
    inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, Lq=196, dist="local", seed=k))
    sets.append(inp)


def fwd_all():
    for inp in sets:
        ops.ms_deform_attn_forward(inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"], 64)


def bwd_all():
    for inp in sets:
        ops.ms_deform_attn_backward(inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"], inp["grad_out"], 64)


REP = 1
for v in (0, 3):
    _lib.set_option("fwd_variant", v)
    print(f"dec spatial fwd (cold)  fwd_variant={v}: {graph_time(fwd_all) / len(sets):6.2f} us", flush=True)
_lib.set_option("fwd_variant", 0)
for chunk in (0, 16, 32):
    _lib.set_option("chunk_pairs", chunk)
    print(f"dec spatial (cold) chunk={chunk}: fwd {graph_time(fwd_all) / len(sets):6.2f} us  bwd (+memset) {graph_time(bwd_all) / len(sets):6.2f} us", flush=True)
_lib.set_option("chunk_pairs", 0)
