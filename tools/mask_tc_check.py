"""Quick correctness + timing probe of the tcgen05 mask kernel (run under `timeout`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops

def nerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())

torch.manual_seed(0)
for (B, Q, K, T, H, W) in [(1, 16, 32, 1, 8, 16), (1, 196, 32, 4, 96, 160), (2, 100, 24, 2, 96, 160), (1, 256, 64, 1, 16, 24), (1, 7, 8, 1, 5, 8), (1, 300 - 44, 32, 2, 33, 24), (1, 196, 32, 4, 160, 288)]:
    coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")).bfloat16()
    proto = torch.randn(B, K, T, H, W, device="cuda").bfloat16()
    want = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    for variant in (1, 3, 2):
        _lib.set_option("mask_variant", variant)
        for od in (torch.float32, torch.bfloat16):
            out = ops.mask_logits_forward(coeff, proto, out_dtype=od)
            torch.cuda.synchronize()
            print(f"B{B} Q{Q} K{K} N{T*H*W} variant {variant} out {od}: nerr {nerr(out, want):.3e}", flush=True)
_lib.set_option("mask_variant", 0)
# timing: device-side (events recorded by the library right around the kernel launch), L2 flushed between launches
flush = torch.empty(128 * 1024 * 1024, device="cuda")
coeff = torch.tanh(torch.randn(1, 196, 32, device="cuda")).bfloat16()
_lib.set_option("profile", 1)
for plane in ((96, 160), (160, 288)):
    proto = torch.randn(1, 32, 4, *plane, device="cuda").bfloat16()
    c32, p32 = coeff.float(), proto.float()
    n = 4 * plane[0] * plane[1]
    for name, variant, cc, pp, od in (("simt f32->f32", 1, c32, p32, torch.float32), ("tc3 f32->f32 (3xTF32)", 0, c32, p32, torch.float32), ("tc1 bf16->f32", 3, coeff, proto, torch.float32),
                                      ("tc1 bf16->bf16", 3, coeff, proto, torch.bfloat16), ("tc2 bf16->f32", 2, coeff, proto, torch.float32),
                                      ("tc2 bf16->bf16", 2, coeff, proto, torch.bfloat16)):
        _lib.set_option("mask_variant", variant)
        ops.mask_logits_forward(cc, pp, out_dtype=od); torch.cuda.synchronize()
        _lib.profile_read(_lib.PROF_MASK_FWD)
        for _ in range(20):
            flush.fill_(1.0)
            ops.mask_logits_forward(cc, pp, out_dtype=od)
        torch.cuda.synchronize()
        ms, cnt = _lib.profile_read(_lib.PROF_MASK_FWD)
        us = ms / cnt * 1e3
        byts = cc.element_size() * (196 * 32 + 32 * n) + (4 if od == torch.float32 else 2) * 196 * n
        print(f"plane {plane} {name}: mean {us:.1f} us -> {byts / us / 1e3:.0f} GB/s, {2 * 196 * 32 * n / us / 1e6:.1f} TFLOP/s", flush=True)
    for name, fn in (("einsum bf16 (cuBLAS)", lambda: torch.einsum("bqm,bmthw->bqthw", coeff, proto)), ("einsum f32 (cuBLAS)", lambda: torch.einsum("bqm,bmthw->bqthw", c32, p32))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print(f"plane {plane} {name}: median {ts[10]:.1f} us (events around the torch call)", flush=True)
# backward (fp32): second-generation SIMT pair vs the first fused kernel
    go = torch.randn(1, 196, 4, *plane, device="cuda")
    for name, variant in (("bwd gen2", 0), ("bwd gen1 fused", 4)):
        _lib.set_option("mask_variant", variant)
        ops.mask_logits_backward(c32, p32, go); torch.cuda.synchronize()
        _lib.profile_read(_lib.PROF_MASK_BWD)
        for _ in range(10):
            flush.fill_(1.0)
            gc, gp = ops.mask_logits_backward(c32, p32, go)
        torch.cuda.synchronize()
        ms, cnt = _lib.profile_read(_lib.PROF_MASK_BWD)
        want_c = torch.einsum("bqthw,bmthw->bqm", go.double(), p32.double())
        want_p = torch.einsum("bqm,bqthw->bmthw", c32.double(), go.double())
        print(f"plane {plane} {name}: {ms / cnt * 1e3:.1f} us  nerr gc {nerr(gc, want_c):.2e} gp {nerr(gp, want_p):.2e}", flush=True)
    _lib.set_option("mask_variant", 0)
_lib.set_option("profile", 0)
