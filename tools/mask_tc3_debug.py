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

# chunked reduction (K > 32) and the tensor-core grad_proto (rows = k, reduction = q, row operand transposed on chip)
for (B, Q, K, N) in [(1, 64, 40, 512), (2, 100, 64, 4096), (1, 196, 128, 8192), (1, 196, 32, 61440), (2, 37, 8, 1000), (1, 196, 32, 7 * 46 * 80)]:
    coeff = torch.tanh(torch.randn(B, Q, K, device="cuda"))
    proto = torch.randn(B, K, 1, 1, N, device="cuda")
    go = torch.randn(B, Q, 1, 1, N, device="cuda")
    want = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    want_gp = torch.einsum("bqm,bqthw->bmthw", coeff.double(), go.double())
    want_gc = torch.einsum("bmthw,bqthw->bqm", proto.double(), go.double())
    _lib.set_option("mask_variant", 0)
    out = ops.mask_logits_forward(coeff, proto)
    gc, gp = ops.mask_logits_backward(coeff, proto, go)
    torch.cuda.synchronize()
    _lib.set_option("mask_variant", 1)
    gc1, gp1 = ops.mask_logits_backward(coeff, proto, go)
    torch.cuda.synchronize()
    _lib.set_option("mask_variant", 5)
    out5 = ops.mask_logits_forward(coeff, proto)
    gc5, gp5 = ops.mask_logits_backward(coeff, proto, go)
    torch.cuda.synchronize()
    _lib.set_option("mask_variant", 0)
    print(f"B{B} Q{Q} K{K} N{N}: fwd tc4 {nerr(out, want):.3e} tc3 {nerr(out5, want):.3e} | grad_proto tc4 {nerr(gp, want_gp):.3e} tc3 {nerr(gp5, want_gp):.3e} "
          f"simt {nerr(gp1, want_gp):.3e} | grad_coeff {nerr(gc, want_gc):.3e}")

# timing of the backward on the bench shape
import time
B, Q, K, N = 1, 196, 32, 4 * 96 * 160
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")); proto = torch.randn(B, K, 4, 96, 160, device="cuda"); go = torch.randn(B, Q, 4, 96, 160, device="cuda")
for variant in (1, 5, 0):
    _lib.set_option("mask_variant", variant)
    for _ in range(3): ops.mask_logits_backward(coeff, proto, go)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.mask_logits_backward(coeff, proto, go)
    b.record(); torch.cuda.synchronize()
    print(f"mask backward variant {variant}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
_lib.set_option("mask_variant", 0)
