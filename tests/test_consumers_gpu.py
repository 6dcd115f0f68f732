"""CUDA kernels of the callers either side of the path (csrc/consumers.cu, through the C ABI) against the committed reference
fixtures, against the numpy oracle on seeded inputs, and -- at the full R50_ovis_360 sizes, where the oracle is too slow --
against the reference's own chain of torch ops executed on the same GPU in fp64."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4          # north star: rel <= 1e-4 in fp32 (normalised max error)


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def nerr(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def pkg():
    import mdqe_cvpr2023_b200 as p
    return p


@pytest.fixture(params=["tensor_core", "simt"])
def match_kernel(request):
    """the matcher costs have two kernels: tcgen05 contraction with the costs formed in the epilogue (csrc/match_cost_tc.cuh; option
    consumer_tc = 2 makes an ineligible shape an error instead of a silent SIMT run) and the SIMT kernel (csrc/consumers.cu)"""
    from mdqe_cvpr2023_b200 import _lib
    _lib.set_option("consumer_tc", 2 if request.param == "tensor_core" else 1)
    yield request.param
    _lib.set_option("consumer_tc", 0)


@pytest.mark.parametrize("name", ["match_cost_K32", "match_cost_K24"])
def test_match_cost_golden(pkg, name, match_kernel):
    d = load(name)
    bce, dice = pkg.mask_match_cost(d["coeff"].cuda(), d["proto"].cuda(), d["targets"].cuda())
    assert nerr(bce, d["cost_bce"]) < TOL and nerr(dice, d["cost_dice"]) < TOL


@pytest.mark.parametrize("Q,K,G,N", [(1, 32, 1, 32), (5, 8, 3, 37), (196, 32, 15, 4 * 24 * 40), (196, 32, 16, 1000), (300, 24, 33, 2048), (255, 32, 4, 64)])
def test_match_cost_vs_oracle(pkg, Q, K, G, N, match_kernel):
    from oracle import consumers_oracle as co
    if match_kernel == "tensor_core" and N % 4:
        pytest.skip("the tensor-core kernel needs 16-byte row strides (auto mode takes the SIMT kernel)")
    g = torch.Generator().manual_seed(Q * 7 + G)
    coeff = torch.tanh(torch.randn(Q, K, generator=g))
    proto = torch.randn(K, N, generator=g)
    tgt = (torch.rand(G, N, generator=g) > 0.6).float()
    tgt[0] = torch.rand(N, generator=g)                    # soft targets are legal too
    bce, dice = pkg.mask_match_cost(coeff.cuda(), proto.cuda(), tgt.cuda())
    want_bce, want_dice = co.match_cost(coeff.numpy(), proto.numpy(), tgt.numpy())
    assert nerr(bce, want_bce) < TOL and nerr(dice, want_dice) < TOL


def test_match_cost_full_size_vs_reference_ops_on_gpu(pkg, match_kernel):
    """R50_ovis_360: Q=196, K=32, plane 4 x 96 x 160; the reference's statements (matcher.py:182, :36-61, :11-28) in fp64 on the GPU."""
    g = torch.Generator().manual_seed(5)
    Q, K, G, T, H, W = 196, 32, 9, 4, 96, 160
    coeff = torch.tanh(torch.randn(Q, K, generator=g)).cuda()
    proto = torch.randn(K, T, H, W, generator=g).cuda()
    tgt = (torch.rand(G, T, H, W, generator=g) > 0.8).float().cuda()
    bce, dice = pkg.mask_match_cost(coeff, proto, tgt)
    x = torch.einsum('qm,mthw->qthw', coeff.double(), proto.double()).flatten(1)
    t = tgt.double().flatten(1)
    pos = torch.nn.functional.binary_cross_entropy_with_logits(x, torch.ones_like(x), reduction="none")
    neg = torch.nn.functional.binary_cross_entropy_with_logits(x, torch.zeros_like(x), reduction="none")
    want_bce = (torch.einsum("nc,mc->nm", pos, t) + torch.einsum("nc,mc->nm", neg, 1 - t)) / x.shape[1]
    s = x.sigmoid()
    want_dice = 1 - (2 * torch.einsum("nc,mc->nm", s, t) + 1) / (s.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)
    assert nerr(bce, want_bce) < TOL and nerr(dice, want_dice) < TOL


def test_match_cost_rejects_cpu_tensors(pkg):
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.mask_match_cost(torch.zeros(2, 32), torch.zeros(32, 8), torch.zeros(1, 8))


@pytest.mark.parametrize("name", ["mask_losses_interinst", "mask_losses_plain"])
def test_mask_losses_golden_with_autograd(pkg, name):
    """criterion.py:440-473 on the reference's own numbers: losses and, through autograd, the gradients of mask_coeff (scattered back
    to the Q rows by the indexing, as in the reference) and proto."""
    d = load(name)
    coeff = d["coeff"].cuda().requires_grad_(True)
    proto = d["proto"].cuda().requires_grad_(True)
    idx = d["src_idx"].cuda()
    ti = d["targets_interinst"].cuda() if int(d["interinst"]) else None
    lm, ld = pkg.mask_losses(coeff[idx], proto, d["targets"].cuda(), ti, float(d["num_masks"]))
    assert abs(float(lm) - float(d["loss_mask"])) < 1e-5 * abs(float(d["loss_mask"]))
    assert abs(float(ld) - float(d["loss_dice"])) < 1e-5 * abs(float(d["loss_dice"]))
    gw = d["grad_weights"].cuda()
    (gw[0] * lm + gw[1] * ld).backward()
    assert nerr(coeff.grad, d["grad_coeff"]) < TOL and nerr(proto.grad, d["grad_proto"]) < TOL


@pytest.mark.parametrize("G,K,N,inter", [(1, 32, 33, True), (9, 32, 4 * 96 * 160, True), (32, 24, 1000, False), (45, 32, 777, True), (0, 32, 64, True)])
def test_mask_losses_vs_oracle(pkg, G, K, N, inter):
    from oracle import consumers_oracle as co
    g = torch.Generator().manual_seed(G * 5 + K)
    coeff = torch.tanh(torch.randn(G, K, generator=g))
    proto = torch.randn(K, N, generator=g)
    tgt = (torch.rand(G, N, generator=g) > 0.7).float()
    ti = (torch.rand(G, N, generator=g) > 0.5).float() if inter else None
    c, p = coeff.cuda().requires_grad_(True), proto.cuda().requires_grad_(True)
    lm, ld = pkg.mask_losses(c, p, tgt.cuda(), ti.cuda() if inter else None, 3.0)
    (0.5 * lm + 2.0 * ld).backward()
    if G == 0:
        assert float(lm) == 0.0 and float(ld) == 0.0 and float(p.grad.abs().max()) == 0.0
        return
    want = co.mask_losses(coeff.numpy(), proto.numpy(), tgt.numpy(), ti.numpy() if inter else None, 3.0, (0.5, 2.0))
    assert abs(float(lm) - want[0]) < 1e-5 * abs(want[0]) and abs(float(ld) - want[1]) < 1e-5 * abs(want[1])
    assert nerr(c.grad, want[2]) < TOL and nerr(p.grad, want[3]) < TOL


@pytest.mark.parametrize("name", ["nms_siou_T4", "nms_siou_T5"])
def test_nms_siou_golden(pkg, name):
    d = load(name)
    assert nerr(pkg.mask_nms_siou(d["mask_pred"].cuda()), d["siou"]) < TOL


@pytest.mark.parametrize("Q,T,H,W", [(1, 1, 2, 2), (7, 2, 9, 13), (100, 4, 96, 160), (127, 5, 24, 40), (130, 3, 12, 20), (300, 6, 12, 20)])
def test_nms_siou_vs_reference_ops_on_gpu(pkg, Q, T, H, W):
    g = torch.Generator().manual_seed(Q + T)
    m = (torch.randn(Q, T, H, W, generator=g) * 2 - 0.3).cuda()
    got = pkg.mask_nms_siou(m)
    nms = m[:, ::2] if T >= 5 else m                                                        # mdqe.py:394
    soft = torch.nn.functional.interpolate(nms, scale_factor=0.5).flatten(1).sigmoid()      # :387
    hard = soft.gt(0.5).double()
    soft = soft.double()
    num = soft @ hard.t()
    want = num / (soft.sum(-1)[:, None] + hard.sum(-1)[None] - num + 1)
    assert nerr(got, want) < TOL


@pytest.mark.parametrize("name", ["track_siou_a", "track_siou_b"])
def test_track_siou_golden(pkg, name):
    d = load(name)
    assert nerr(pkg.mask_track_siou(d["saved_masks"].cuda(), d["input_masks"].cuda()), d["siou"]) < TOL


@pytest.mark.parametrize("Ns,Ni,T,H,W", [(1, 1, 1, 1, 1), (20, 35, 4, 96, 160), (130, 5, 2, 12, 20), (4, 260, 1, 7, 9)])
def test_track_siou_vs_oracle(pkg, Ns, Ni, T, H, W):
    from oracle import consumers_oracle as co
    g = torch.Generator().manual_seed(Ns * 3 + Ni)
    saved = torch.rand(Ns, T, H, W, generator=g)
    inp = torch.rand(Ni, T, H, W, generator=g)
    inp[0] = 0.1                                            # an empty mask: the whole column is 0
    got = pkg.mask_track_siou(saved.cuda(), inp.cuda())
    want = co.track_siou(saved.numpy(), inp.numpy())
    assert nerr(got, want) < TOL
    assert float(got[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["aligned_bilinear_f4", "aligned_bilinear_f2"])
def test_aligned_bilinear_golden(pkg, name):
    d = load(name)
    f = int(d["factor"])
    assert nerr(pkg.aligned_bilinear(d["x"].cuda(), f), d["up"]) < 1e-6
    assert nerr(pkg.aligned_bilinear(d["x"].cuda(), f, sigmoid=True), d["up_sigmoid"]) < 1e-6


@pytest.mark.parametrize("shape,f", [((1, 1, 1, 1), 4), ((2, 3, 7, 5), 3), ((10, 4, 96, 160), 4), ((3, 2, 5, 9), 1), ((1, 2, 6, 7), 2)])
def test_aligned_bilinear_vs_oracle(pkg, shape, f):
    from oracle import consumers_oracle as co
    g = torch.Generator().manual_seed(sum(shape) + f)
    x = torch.randn(*shape, generator=g) * 4
    got = pkg.aligned_bilinear(x.cuda(), f)
    want = co.aligned_bilinear(x.numpy().astype(np.float64), f)
    assert tuple(got.shape) == want.shape
    assert nerr(got, want) < 1e-6
    # size-independent property: a constant image stays constant, a horizontal ramp stays linear in the interior
    c = pkg.aligned_bilinear(torch.full(shape, 2.5).cuda(), f)
    assert float((c - 2.5).abs().max()) == 0.0


def test_query_init_golden_and_autograd(pkg):
    d = load("query_init_f64")
    feat = d["feat"].float().cuda().requires_grad_(True)
    coords = d["coords"].float().cuda().requires_grad_(True)
    out = pkg.query_init_sample(feat, d["shapes"].cuda(), d["level_start"].cuda(), coords)
    assert nerr(out, d["out"]) < 1e-5
    out.backward(d["grad_out"].float().cuda())
    assert nerr(feat.grad, d["grad_feat"]) < 1e-5
    assert nerr(coords.grad, d["grad_coords"]) < TOL


def test_query_init_r50_shape_vs_grid_sample_on_gpu(pkg):
    """B*T = 4 frames, 196 query points, C = 256, the four R50_ovis_360 levels; transformer_dec.py:172-179 in fp64 on the GPU."""
    g = torch.Generator().manual_seed(3)
    shapes_list = [(48, 80), (24, 40), (12, 20), (6, 10)]
    S = sum(h * w for h, w in shapes_list)
    starts = [0]
    for h, w in shapes_list:
        starts.append(starts[-1] + h * w)
    feat = torch.randn(4, S, 256, generator=g).cuda().requires_grad_(True)
    coords = (torch.rand(4, 196, 2, generator=g) * 1.04 - 0.02).cuda().requires_grad_(True)
    go = torch.randn(4, 196, 256, generator=g).cuda()
    out = pkg.query_init_sample(feat, torch.tensor(shapes_list).cuda(), torch.tensor(starts[:-1]).cuda(), coords)
    out.backward(go)
    f64 = feat.detach().double().requires_grad_(True)
    c64 = coords.detach().double().requires_grad_(True)
    grid = (2 * c64 - 1).view(4, 14, 14, 2)
    ref = []
    for l, (H_l, W_l) in enumerate(shapes_list):
        ref.append(torch.nn.functional.grid_sample(f64[:, starts[l]:starts[l + 1]].transpose(1, 2).reshape(4, 256, H_l, W_l), grid, mode='bilinear',
                                                   padding_mode="border", align_corners=False))
    ref = torch.stack(ref).mean(0).flatten(2).transpose(1, 2)
    ref.backward(go.double())
    assert nerr(out, ref) < 1e-5
    assert nerr(feat.grad, f64.grad) < 1e-5
    assert nerr(coords.grad, c64.grad) < TOL
