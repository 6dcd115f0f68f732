"""Clock-stamp timeline of CTA 0 of mask_grad_coeff_tc_kernel (option mask_debug=1), bench shape."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
B, Q, K = 1, 196, 32
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")); proto = torch.randn(B, K, 4, 96, 160, device="cuda"); go = torch.randn(B, Q, 4, 96, 160, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ops.mask_logits_backward(coeff, proto, go, need_proto=False); torch.cuda.synchronize()
flush.zero_()
_lib.set_option("mask_debug", 1)
ops.mask_logits_backward(coeff, proto, go, need_proto=False); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 80)()
_lib.check(_lib.load().msda_debug_read(buf), "debug")
_lib.set_option("mask_debug", 0)
t = [list(buf[r * 16:(r + 1) * 16]) for r in range(4)]
t0 = min(x for x in t[0] if x)
print("cycles relative to the first TMA issue (CTA 0), chunks 0..15:")
for name, row in zip(("tma_issued", "landed", "split_done", "mma_issue"), t):
    print("  %-11s" % name, [x - t0 if x else None for x in row])
