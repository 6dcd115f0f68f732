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


inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, Lq=196, dist="local", seed=0))
a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
for key, vals in (("fwd_variant", (0, 3)),):
    for v in vals:
        _lib.set_option(key, v)
        print(f"dec spatial fwd  {key}={v}: {graph_time(lambda: ops.ms_deform_attn_forward(*a, 64)):6.2f} us", flush=True)
    _lib.set_option(key, 0)
for chunk in (0, 16, 32):
    _lib.set_option("chunk_pairs", chunk)
    f = graph_time(lambda: ops.ms_deform_attn_forward(*a, 64))
    b = graph_time(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64))
    print(f"dec spatial chunk={chunk}: fwd {f:6.2f} us  bwd (+memset) {b:6.2f} us", flush=True)
_lib.set_option("chunk_pairs", 0)
