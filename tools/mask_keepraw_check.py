import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
def nerr(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
torch.manual_seed(0)
B, Q, K, N = 1, 196, 32, 4 * 96 * 160
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")); proto = torch.randn(B, K, 4, 96, 160, device="cuda"); go = torch.randn(B, Q, 4, 96, 160, device="cuda")
want_gc = torch.einsum("bmthw,bqthw->bqm", proto.double(), go.double())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dbg in (2, 0, 2, 0):
    _lib.set_option("mask_debug", dbg)
    gc, _ = ops.mask_logits_backward(coeff, proto, go, need_proto=False)
    torch.cuda.synchronize()
    tot = 0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.mask_logits_backward(coeff, proto, go, need_proto=False); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print(f"mask_debug {dbg} (2 = rewrite hi in place, 0 = keep the raw tile as hi): grad_coeff nerr {nerr(gc, want_gc):.3e}  {tot / 10 * 1e3:.1f} us")
_lib.set_option("mask_debug", 0)
