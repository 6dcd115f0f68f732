"""Tensor-core Linear (3xTF32 GEMM, csrc/gemm3x.cuh) against F.linear in fp64: forward with bias / row mask, dgrad, wgrad.
Tolerance: 1e-4 normalised as for every fp32 kernel of the path (asserted 2e-5)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import nerr

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["stream_k", "tiles"], autouse=True)
def work_distribution(request):
    """Both work distributions of the GEMM: contiguous (tile, chunk) ranges per CTA with tiles combined in the output (default) and
    whole tiles dealt round-robin (option gemm_stream_k = 0)."""
    from mdqe_cvpr2023_b200 import _lib
    _lib.set_option("gemm_stream_k", 2 if request.param == "stream_k" else 0)              # 2 = on every shape, 1 (default) = where it pays
    yield request.param
    _lib.set_option("gemm_stream_k", 1)


@pytest.mark.parametrize("rows,in_f,out_f", [(4 * 5100, 256, 256), (4 * 5100, 256, 128), (784, 256, 256), (3 * 5100, 192, 192),
                                             (3 * 5100, 192, 96), (1, 256, 256), (130, 36, 20), (257, 8, 300), (1000, 260, 4)])
def test_tc_linear_vs_fp64(rows, in_f, out_f):
    from mdqe_cvpr2023_b200 import tc_linear
    g = torch.Generator(device="cuda").manual_seed(rows + in_f)
    x = torch.randn(rows, in_f, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(out_f, in_f, device="cuda", generator=g) / in_f ** 0.5).requires_grad_(True)
    b = torch.randn(out_f, device="cuda", generator=g, requires_grad=True)
    mask = torch.rand(rows, device="cuda", generator=g) < 0.1
    gy = torch.randn(rows, out_f, device="cuda", generator=g)
    y = tc_linear(x, w, b, mask)
    y.backward(gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = F.linear(xd, wd, bd).masked_fill(mask[:, None], 0.0)
    yd.backward(gy.double())
    assert nerr(y, yd) < 2e-5
    assert nerr(x.grad, xd.grad) < 2e-5
    assert nerr(w.grad, wd.grad) < 2e-5
    assert nerr(b.grad, bd.grad) < 2e-5
    # no bias, no mask, leading batch dims
    x3 = x.detach().view(1, rows, in_f) if rows % 2 else x.detach().view(2, rows // 2, in_f)
    y2 = tc_linear(x3, w.detach())
    assert y2.shape == x3.shape[:-1] + (out_f,)
    assert nerr(y2.reshape(rows, out_f), F.linear(x.detach().double(), w.detach().double())) < 2e-5


@pytest.mark.parametrize("rows,in_f,out_f", [(4 * 5100, 256, 256), (784, 256, 128), (130, 36, 20), (1, 8, 4), (0, 8, 4)])
def test_backward_with_bias_in_one_call(rows, in_f, out_f):
    """tc_linear_backward_bias: the bias gradient out of the weight-gradient GEMM (column sums taken by the staging warps), and the
    fallback to the bias kernel when no weight gradient is asked for."""
    from mdqe_cvpr2023_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + out_f)
    x = torch.randn(rows, in_f, device="cuda", generator=g)
    w = torch.randn(out_f, in_f, device="cuda", generator=g) / in_f ** 0.5
    gy = torch.randn(rows, out_f, device="cuda", generator=g) + 0.25
    want_b = gy.double().sum(0)
    want_w = gy.double().t() @ x.double()
    want_x = gy.double() @ w.double()
    for need_x, need_w in ((True, True), (False, True), (True, False), (False, False)):
        gx, gw, gb = ops.tc_linear_backward(gy, x, w, need_x=need_x, need_weight=need_w, need_bias=True)
        assert (gx is None) == (not need_x) and (gw is None) == (not need_w)
        scale = max(float(want_b.abs().max()), 1e-30)
        assert float((gb.double() - want_b).abs().max()) / scale < 2e-5 or rows == 0
        if rows == 0:
            assert float(gb.abs().max()) == 0.0
            continue
        if need_x:
            assert nerr(gx, want_x) < 2e-5
        if need_w:
            assert nerr(gw, want_w) < 2e-5


def test_repeated_launches_leave_the_tile_flags_clean():
    """The stream-K form combines split tiles through per-tile flags that every launch must leave zero: many launches in a row, on two
    streams, inside a CUDA graph, all against the same fp64 answer."""
    from mdqe_cvpr2023_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    cases = []
    for rows, in_f, out_f in ((4 * 5100, 256, 256), (784, 256, 128), (300, 512, 96)):
        x = torch.randn(rows, in_f, device="cuda", generator=g)
        w = torch.randn(out_f, in_f, device="cuda", generator=g) / in_f ** 0.5
        b = torch.randn(out_f, device="cuda", generator=g)
        cases.append((x, w, b, F.linear(x.double(), w.double(), b.double())))
    for _ in range(40):
        for x, w, b, want in cases:
            assert nerr(ops.tc_linear_forward(x, w, b), want) < 2e-5
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for i in range(20):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                x, w, b, want = cases[i % 3]
                outs.append((ops.tc_linear_forward(x, w, b), want))
    torch.cuda.synchronize()
    for y, want in outs:
        assert nerr(y, want) < 2e-5
    x, w, b, want = cases[0]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.tc_linear_forward(x, w, b)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = ops.tc_linear_forward(x, w, b)
        gx, _ = ops.tc_linear_backward(y, x, w, True, False)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    assert nerr(y, want) < 2e-5
    assert nerr(gx, want @ w.double()) < 2e-5
    from mdqe_cvpr2023_b200 import _lib
    assert _lib.load().msda_gemm_flag_timeouts() == 0, "a stream-K stretch gave up waiting for its tile's flag"


def test_tc_linear_rejects_unsupported():
    from mdqe_cvpr2023_b200 import ops
    x = torch.randn(8, 6, device="cuda")
    w = torch.randn(4, 6, device="cuda")
    assert not ops.linear_supported(x, w)
    with pytest.raises(RuntimeError, match="multiples of 4"):
        ops.tc_linear_forward(x, w)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.tc_linear_forward(torch.randn(8, 8), torch.randn(8, 8))
