"""Per-kernel timing of the mask-head backward on the bench shape (device events, L2 flushed between calls)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops

B, Q, K = 1, 196, 32
T, H, W = 4, 96, 160
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda"))
proto = torch.randn(B, K, T, H, W, device="cuda")
go = torch.randn(B, Q, T, H, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, reps=10):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3

for variant in (1, 0):
    _lib.set_option("mask_variant", variant)
    print(f"variant {variant}: fwd {timed(lambda: ops.mask_logits_forward(coeff, proto)):.1f} us | "
          f"grad_coeff only {timed(lambda: ops.mask_logits_backward(coeff, proto, go, need_proto=False)):.1f} us | "
          f"grad_proto only {timed(lambda: ops.mask_logits_backward(coeff, proto, go, need_coeff=False)):.1f} us | "
          f"both {timed(lambda: ops.mask_logits_backward(coeff, proto, go)):.1f} us")
_lib.set_option("mask_variant", 0)
