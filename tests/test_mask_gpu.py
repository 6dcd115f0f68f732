"""Mask contraction (coeff x proto) on the GPU against the einsum fixtures and torch.einsum (config 5
sweep, reduced).  fp32 tolerance 1e-4 normalised (asserted 2e-5), bf16 2e-2."""
import itertools

import pytest
import torch

from tests.helpers import load_golden, nerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_options():
    from mdqe_cvpr2023_b200 import _lib
    yield
    _lib.set_option("mask_variant", 0)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", ["mask_einsum_K32", "mask_einsum_K24"])
def test_mask_golden(name, variant):
    from mdqe_cvpr2023_b200 import _lib, mask_logits
    _lib.set_option("mask_variant", variant)
    z = load_golden(name)
    coeff = torch.from_numpy(z["coeff"]).cuda().requires_grad_(True)
    proto = torch.from_numpy(z["proto"]).cuda().requires_grad_(True)
    out = mask_logits(coeff, proto)
    assert tuple(out.shape) == tuple(z["out"].shape)
    assert nerr(out, z["out"]) < 2e-5
    out.backward(torch.from_numpy(z["grad_out"]).cuda())
    assert nerr(coeff.grad, z["grad_coeff"]) < 2e-5
    assert nerr(proto.grad, z["grad_proto"]) < 2e-5


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("Q,T,plane,K", [(196, 4, (96, 160), 32), (100, 2, (96, 160), 24), (300, 2, (160, 288), 32),
                                         (196, 3, (96, 160), 24), (7, 1, (5, 9), 32), (130, 2, (33, 17), 40), (256, 2, (160, 288), 64)])
def test_mask_sweep_vs_einsum(Q, T, plane, K, variant):
    from mdqe_cvpr2023_b200 import _lib, ops
    _lib.set_option("mask_variant", variant)
    g = torch.Generator(device="cuda").manual_seed(Q + T)
    coeff = torch.tanh(torch.randn(1, Q, K, device="cuda", generator=g))
    proto = torch.randn(1, K, T, *plane, device="cuda", generator=g)
    want = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    out = ops.mask_logits_forward(coeff, proto)
    assert nerr(out, want) < 2e-5
    out16 = ops.mask_logits_forward(coeff.bfloat16(), proto.bfloat16())
    want16 = torch.einsum("bqm,bmthw->bqthw", coeff.bfloat16().double(), proto.bfloat16().double())
    assert out16.dtype == torch.bfloat16 and nerr(out16.float(), want16) < 2e-2
    out_mixed = ops.mask_logits_forward(coeff, proto, out_dtype=torch.bfloat16)
    assert nerr(out_mixed.float(), want) < 2e-2


CONFIG5 = [(Q, T, plane, K) for Q in (100, 196, 300) for T in (2, 4, 8) for plane in ((96, 160), (160, 288)) for K in (32, 24)
           if not (T == 8 and plane == (160, 288) and Q == 300)] + [(300, 8, (160, 288), 32)]


@pytest.mark.parametrize("Q,T,plane,K", CONFIG5[::3] + [(7, 1, (5, 9), 32), (130, 2, (33, 17), 40), (64, 2, (12, 20), 20)])
@pytest.mark.parametrize("variant", [0, 1])
def test_mask_fp16_vs_fp16_einsum(Q, T, plane, K, variant):
    """The reference's evaluation dtype (autocast: train_net.py:207-208 -> mdqe/mdqe.py:384 runs in fp16): fp16 operands,
    fp16 result, against torch.einsum on the same fp16 tensors (<= 2e-2) and against the exact product of the rounded
    operands; BASELINE config 5 sweep (every third point) plus odd sizes that take the SIMT kernel."""
    from mdqe_cvpr2023_b200 import _lib, ops
    _lib.set_option("mask_variant", variant)
    g = torch.Generator(device="cuda").manual_seed(Q * 7 + T)
    coeff = torch.tanh(torch.randn(1, Q, K, device="cuda", generator=g)).half()
    proto = torch.randn(1, K, T, *plane, device="cuda", generator=g).half()
    out = ops.mask_logits_forward(coeff, proto)
    assert out.dtype == torch.float16 and tuple(out.shape) == (1, Q, T) + plane
    ref16 = torch.einsum("bqm,bmthw->bqthw", coeff, proto)
    exact = torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())
    assert nerr(out.float(), ref16.float()) < 2e-2
    assert nerr(out.float(), exact) < 2e-3                                     # fp32 accumulation, one rounding to fp16
    out32 = ops.mask_logits_forward(coeff, proto, out_dtype=torch.float32)
    assert out32.dtype == torch.float32 and nerr(out32, exact) < 1e-5
    with pytest.raises(RuntimeError):
        ops.mask_logits_forward(coeff, proto, out_dtype=torch.bfloat16)


def test_mask_logits_under_autocast_matches_einsum_semantics():
    """`mask_logits` is a drop-in for the einsum at the reference's call sites under the reference's own eval setting."""
    from mdqe_cvpr2023_b200 import mask_logits
    g = torch.Generator(device="cuda").manual_seed(5)
    coeff = torch.tanh(torch.randn(2, 196, 32, device="cuda", generator=g)).requires_grad_(True)
    proto = torch.randn(2, 32, 4, 24, 40, device="cuda", generator=g).requires_grad_(True)
    for dt in (torch.float16, torch.bfloat16):
        with torch.autocast("cuda", dtype=dt):
            want = torch.einsum("bqm,bmthw->bqthw", coeff, proto)
            got = mask_logits(coeff, proto)
            got1 = mask_logits(coeff[0], proto[0])                              # the unbatched form of mdqe/mdqe.py:384
        assert got.dtype == want.dtype == dt and got1.dtype == dt
        assert nerr(got.float(), want.float()) < 2e-2 and nerr(got1.float(), want[0].float()) < 2e-2
        gw = torch.autograd.grad(want.float().sum(), (coeff, proto), retain_graph=True)
        gg = torch.autograd.grad(got.float().sum(), (coeff, proto))
        for a, b in zip(gg, gw):
            assert a.dtype == b.dtype == torch.float32 and nerr(a, b) < 2e-2
    out = mask_logits(coeff, proto)                                             # outside autocast: fp32 as before
    assert out.dtype == torch.float32 and nerr(out, torch.einsum("bqm,bmthw->bqthw", coeff.double(), proto.double())) < 2e-5


def test_mask_unbatched_form_and_empty():
    from mdqe_cvpr2023_b200 import mask_logits, ops
    coeff = torch.randn(5, 32, device="cuda")
    proto = torch.randn(32, 2, 6, 8, device="cuda")
    assert nerr(mask_logits(coeff, proto), torch.einsum("qm,mthw->qthw", coeff, proto)) < 2e-5   # mdqe/mdqe.py:384
    out = ops.mask_logits_forward(torch.zeros(1, 0, 32, device="cuda"), torch.zeros(1, 32, 1, 4, 4, device="cuda"))
    assert tuple(out.shape) == (1, 0, 1, 4, 4)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("B,Q,K,N", [(1, 196, 32, 4 * 24 * 40), (2, 100, 24, 1000), (1, 300, 32, 2048 + 36), (2, 37, 8, 516),
                                     (1, 64, 40, 512), (1, 196, 128, 4096), (1, 520, 64, 1024), (1, 9, 32, 30)])
def test_mask_backward_sweep(B, Q, K, N, variant):
    """grad_coeff / grad_proto of the contraction against fp64 einsum: tensor-core kernels (variant 0: MN-major operands straight
    from TMA + TMEM-resident grad_coeff) and the SIMT kernels (1).  Covers several
    batch items, K above one 32-wide reduction chunk, more than 256 query rows (two row blocks), column counts that are not a
    multiple of the 32-column chunk / 128-column tile, and N % 4 != 0 (tensor-core path ineligible -> SIMT)."""
    from mdqe_cvpr2023_b200 import _lib, ops
    _lib.set_option("mask_variant", variant)
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Q + K)
    coeff = torch.tanh(torch.randn(B, Q, K, device="cuda", generator=g))
    proto = torch.randn(B, K, 1, 1, N, device="cuda", generator=g)
    go = torch.randn(B, Q, 1, 1, N, device="cuda", generator=g)
    gc, gp = ops.mask_logits_backward(coeff, proto, go)
    want_gp = torch.einsum("bqm,bqthw->bmthw", coeff.double(), go.double())
    want_gc = torch.einsum("bmthw,bqthw->bqm", proto.double(), go.double())
    assert nerr(gc, want_gc) < 2e-5 and nerr(gp, want_gp) < 2e-5
    gc_only, none = ops.mask_logits_backward(coeff, proto, go, need_proto=False)
    assert none is None and nerr(gc_only, want_gc) < 2e-5
    none, gp_only = ops.mask_logits_backward(coeff, proto, go, need_coeff=False)
    assert none is None and nerr(gp_only, want_gp) < 2e-5
