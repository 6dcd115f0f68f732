import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
def nerr(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
torch.manual_seed(0)
for (B, Q, K, N) in [(1, 32, 32, 128), (1, 32, 8, 128), (1, 64, 32, 256), (1, 196, 32, 61440), (1, 100, 24, 30720)]:
    coeff = torch.tanh(torch.randn(B, Q, K, device="cuda"))
    proto = torch.randn(B, K, 1, 1, N, device="cuda")
    want = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    _lib.set_option("mask_variant", 0)
    out = ops.mask_logits_forward(coeff, proto); torch.cuda.synchronize()
    # what single-pass TF32 (hi*hi only) would give, and hi*hi+hi*lo+lo*hi
    def trunc(t): return (t.view(torch.int32) & -8192).view(torch.float32)
    ch, ph = trunc(coeff), trunc(proto)
    cl, pl = coeff - ch, proto - ph
    hh = torch.einsum("bqm,bmthw->bqthw", ch.double(), ph.double())
    three = hh + torch.einsum("bqm,bmthw->bqthw", ch.double(), pl.double()) + torch.einsum("bqm,bmthw->bqthw", cl.double(), ph.double())
    print(f"Q{Q} K{K} N{N}: nerr vs exact {nerr(out, want):.3e} | vs hi*hi only {nerr(out, hh):.3e} | vs 3-term {nerr(out, three):.3e} | 3-term vs exact {nerr(three, want):.3e}")
    if Q == 32 and K == 32:
        # structure probes: which (q, n) entries are wrong?
        err = (out.double() - want).abs().view(Q, N)
        print("   max err per 16-row q block:", [f"{err[i:i+16].max().item():.2e}" for i in range(0, Q, 16)])
        print("   max err per 32-col n block:", [f"{err[:, i:i+32].max().item():.2e}" for i in range(0, N, 32)])
