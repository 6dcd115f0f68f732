#!/usr/bin/env python
"""A/B of the plane operand through tensor memory in the 3xTF32 mask kernels (option mask_a_tmem: 1 = never, 0 = auto, 2 = always):
forward, grad_proto alone, full backward at the BASELINE mask shapes.  CUDA-graph replays (10 calls per graph)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402

def timed(fn, reps=20):
    """10 calls per CUDA graph, replayed: GPU time per call without host dispatch (inputs larger than L2 at every shape but the smallest)"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (10 * reps)


for Q, T, plane, K in ((196, 4, (96, 160), 32), (196, 4, (160, 288), 32), (300, 8, (96, 160), 32), (196, 3, (96, 160), 24)):
    n = T * plane[0] * plane[1]
    coeff = torch.tanh(torch.randn(1, Q, K, device="cuda"))
    proto = torch.randn(1, K, T, *plane, device="cuda")
    go = torch.randn(1, Q, T, *plane, device="cuda")
    want_f = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    want_p = torch.einsum("bqm,bqthw->bmthw", coeff.double(), go.double())
    for opt in (1, 0, 2):
        _lib.set_option("mask_a_tmem", opt)
        out = ops.mask_logits_forward(coeff, proto)
        gp = ops.mask_logits_backward(coeff, proto, go, need_coeff=False)[1]
        ef = float((out.double() - want_f).abs().max() / want_f.abs().max())
        ep = float((gp.double() - want_p).abs().max() / want_p.abs().max())
        print("Q%d K%d N%d  mask_a_tmem=%d  fwd %6.1f us (err %.1e)  grad_proto %6.1f us (err %.1e)  backward %6.1f us" % (
            Q, K, n, opt, timed(lambda: ops.mask_logits_forward(coeff, proto)), ef,
            timed(lambda: ops.mask_logits_backward(coeff, proto, go, need_coeff=False)), ep,
            timed(lambda: ops.mask_logits_backward(coeff, proto, go))), flush=True)
    _lib.set_option("mask_a_tmem", 0)
