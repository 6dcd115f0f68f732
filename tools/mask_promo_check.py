"""Experiment: TMA L2 promotion 128B vs 256B (mask_debug=3) on the fp32 mask kernels, bench shape, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
B, Q, K = 1, 196, 32
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")); proto = torch.randn(B, K, 4, 96, 160, device="cuda"); go = torch.randn(B, Q, 4, 96, 160, device="cuda")
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
for dbg in (0, 3, 0, 3):
    _lib.set_option("mask_debug", dbg)
    print(f"mask_debug {dbg}: fwd {timed(lambda: ops.mask_logits_forward(coeff, proto)):.1f} | grad_coeff {timed(lambda: ops.mask_logits_backward(coeff, proto, go, need_proto=False)):.1f} | "
          f"grad_proto {timed(lambda: ops.mask_logits_backward(coeff, proto, go, need_coeff=False)):.1f} us")
_lib.set_option("mask_debug", 0)
