import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
coeff = torch.tanh(torch.randn(1, 196, 32, device="cuda")).bfloat16()
proto = torch.randn(1, 32, 4, 96, 160, device="cuda").bfloat16()
_lib.set_option("mask_variant", 2)
for od in (torch.float32, torch.bfloat16):
    ops.mask_logits_forward(coeff, proto, out_dtype=od); torch.cuda.synchronize()
    _lib.set_option("mask_debug", 1)
    ops.mask_logits_forward(coeff, proto, out_dtype=od); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 80)()
    _lib.check(_lib.load().msda_debug_read(buf), "debug")
    _lib.set_option("mask_debug", 0)
    t = [list(buf[r * 16:(r + 1) * 16]) for r in range(5)]
    t0 = min(x for x in t[0] if x)
    print("out", od, "cycles relative to first TMA issue (CTA 0):")
    for name, row in zip(("tma_issued", "mma_wait", "operands_in", "epi_start", "epi_done"), t):
        print("  %-12s" % name, [x - t0 if x else None for x in row[:8]])
