"""Stack-level parity on the GPU (build-plan test t4): the callers of the op -- the encoder layer
(/root/reference/mdqe/models/transformer_enc.py:100-110: self-attention on the flattened pyramid + residual + LayerNorm + FFN) and
the decoder layer's two cross-attentions (transformer_dec.py:340-346 frame-level `forward_ca_box`, :361-395 clip-level
`forward_ca_inst` in temporal mode) -- restated here with this package's `MSDeformAttn`, two layers each, run

  * on cuda:0 through the CUDA path exactly as shipped (fused prologue, tensor-core Linear layers, grouped temporal launch,
    side-stream zero-fill), and
  * on the CPU with the same weights, torch Linear layers and the plain-C oracle as the Function,

and compared on outputs, input gradients and every parameter gradient.  The reference tree is not available on the GPU box, so its
layer code cannot be imported there; tests/test_reference_stack_cpu.py runs the reference's own classes on this module (CPU) and
the two tests meet at `mdqe_cvpr2023_b200.MSDeformAttn`."""
import copy

import pytest
import torch
from torch import nn

from tests.helpers import OracleMSDAFunction, nerr

pytestmark = pytest.mark.gpu

DIM, HEADS, POINTS, FFN = 256, 8, 4, 512
PYRAMID = [(24, 40), (12, 20), (6, 10), (3, 5)]


class EncLayer(nn.Module):
    """transformer_enc.py:83-110: src2 = self_attn(src + pos, ref, src, shapes, padding) -> residual -> norm -> FFN -> norm"""

    def __init__(self, attn_cls):
        super().__init__()
        self.self_attn = attn_cls(DIM, len(PYRAMID), HEADS, POINTS)
        self.norm1, self.norm2 = nn.LayerNorm(DIM), nn.LayerNorm(DIM)
        self.linear1, self.linear2 = nn.Linear(DIM, FFN), nn.Linear(FFN, DIM)

    def forward(self, src, pos, ref, shapes, padding):
        src = self.norm1(src + self.self_attn(src + pos, ref, src, shapes, padding))
        return self.norm2(src + self.linear2(torch.relu(self.linear1(src))))


class DecLayer(nn.Module):
    """transformer_dec.py:340-346 + :361-395 without the query self-attention: frame-level cross-attention on [B*T] frames with
    box reference points, then clip-level cross-attention where the T frames of a clip play the role of levels."""

    def __init__(self, attn_cls, T):
        super().__init__()
        self.T = T
        self.cross_attn_box = attn_cls(DIM, len(PYRAMID), HEADS, POINTS, pred_offsets=False, mode="spatial")
        self.temp_attn_inst = attn_cls(DIM, len(PYRAMID), HEADS, POINTS, n_frames=T, pred_offsets=False, mode="temporal")
        self.norm_box, self.norm_inst = nn.LayerNorm(DIM), nn.LayerNorm(DIM)

    def forward(self, q_box, q_inst, pos_box, pos_inst, boxes, inst_boxes, memory, shapes, padding):
        BT, S, _ = memory.shape
        q_box = self.norm_box(q_box + self.cross_attn_box(q_box + pos_box, boxes, memory, shapes, padding))
        mem_clip = memory.view(BT // self.T, self.T, S, DIM)
        pad_clip = padding.view(BT // self.T, self.T, S) if padding is not None else None
        q_inst = self.norm_inst(q_inst + self.temp_attn_inst(q_inst + pos_inst, inst_boxes, mem_clip, shapes, pad_clip))
        return q_box, q_inst


class Stack(nn.Module):
    def __init__(self, attn_cls, T, n_layers=2):
        super().__init__()
        self.enc = nn.ModuleList(EncLayer(attn_cls) for _ in range(n_layers))
        self.dec = nn.ModuleList(DecLayer(attn_cls, T) for _ in range(n_layers))

    def forward(self, src, pos, enc_ref, shapes, padding, q_box, q_inst, pos_box, pos_inst, boxes, inst_boxes):
        for layer in self.enc:
            src = layer(src, pos, enc_ref, shapes, padding)
        for layer in self.dec:
            q_box, q_inst = layer(q_box, q_inst, pos_box, pos_inst, boxes, inst_boxes, src, shapes, padding)
        return src, q_box, q_inst


def _inputs(B, T, Q, g):
    S = sum(h * w for h, w in PYRAMID)
    pts = []
    for H, W in PYRAMID:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
    pix = torch.cat(pts)
    r = lambda *s: torch.randn(*s, generator=g)
    boxes = torch.cat([torch.rand(B * T, Q, 2, generator=g), torch.rand(B * T, Q, 2, generator=g) * 0.2 + 0.05], -1)
    return dict(src=r(B * T, S, DIM), pos=0.1 * r(B * T, S, DIM),
                enc_ref=torch.cat([pix, torch.full_like(pix, 0.1)], -1).unsqueeze(0).expand(B * T, S, 4).contiguous(),
                shapes=torch.tensor(PYRAMID), padding=torch.rand(B * T, S, generator=g) < 0.05,
                q_box=r(B * T, Q, DIM), q_inst=r(B, Q, DIM), pos_box=0.1 * r(B * T, Q, DIM), pos_inst=0.1 * r(B, Q, DIM),
                boxes=boxes, inst_boxes=boxes[::T].contiguous())


def _run_gpu(stack_cpu, inp, w, grad_names, tc_linear):
    import mdqe_cvpr2023_b200.modules as M
    stack = copy.deepcopy(stack_cpu).cuda()
    for m in stack.modules():
        if isinstance(m, M.MSDeformAttn):
            m.tc_linear = tc_linear
    dev = {k: (v.cuda() if v is not None else None) for k, v in inp.items()}
    for k in grad_names:
        dev[k].requires_grad_(True)
    outs = stack(**dev)
    sum((o * wi.cuda()).sum() for o, wi in zip(outs, w)).backward()
    torch.cuda.synchronize()
    return outs, {k: dev[k].grad for k in grad_names}, {k: p.grad for k, p in stack.named_parameters()}


@pytest.mark.parametrize("padding", [True, False])
def test_encoder_decoder_stack_cuda_vs_oracle(padding, monkeypatch):
    import mdqe_cvpr2023_b200.modules as M
    from mdqe_cvpr2023_b200 import _lib
    B, T, Q = 2, 3, 50
    torch.manual_seed(0)
    stack_cpu = Stack(M.MSDeformAttn, T)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                                          # move the weights off their zero / grid initialisation
        for p in stack_cpu.parameters():
            p.add_(0.02 * torch.randn(p.shape, generator=g))
    inp = _inputs(B, T, Q, g)
    if not padding:
        inp["padding"] = None
    grad_names = ("src", "q_box", "q_inst")
    w = [torch.randn(s, generator=g) for s in ((B * T, sum(h * w for h, w in PYRAMID), DIM), (B * T, Q, DIM), (B, Q, DIM))]

    # ---- device under test: the CUDA path (fused prologue, grouped temporal launch, side-stream zero-fill), once with the
    # Linear layers on torch (sampling locations then agree with the CPU run to the last bits) and once as shipped (3xTF32
    # tensor-core Linear layers)
    _lib.launch_count_reset()
    exact = _run_gpu(stack_cpu, inp, w, grad_names, tc_linear=False)
    shipped = _run_gpu(stack_cpu, inp, w, grad_names, tc_linear=True)
    assert _lib.launch_count() >= 2 * (2 * 3 * 2 * 2), "the CUDA kernels did not run"      # sampler fwd+bwd of 6 modules, twice

    # ---- truth: same weights on the CPU, torch Linear layers + the plain-C oracle as the Function, reference op sequence
    monkeypatch.setattr(M, "MSDeformAttnFunction", OracleMSDAFunction)
    for m in stack_cpu.modules():
        if isinstance(m, M.MSDeformAttn):
            m.fused_prologue, m.tc_linear = False, False
    monkeypatch.setattr(M.ops, "grouped_supported", lambda *a: False)            # per-level Function calls, like the reference
    cpu = {k: (v.clone() if v is not None else None) for k, v in inp.items()}
    for k in grad_names:
        cpu[k].requires_grad_(True)
    outs_cpu = stack_cpu(**cpu)
    sum((o * wi).sum() for o, wi in zip(outs_cpu, w)).backward()
    pg = {k: p.grad for k, p in stack_cpu.named_parameters()}

    def close(a, b, what, max_l2, max_frac):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        rel_l2 = float((a - b).norm() / b.norm())
        frac = float(((a - b).abs() > 1e-3 * b.abs().max()).double().mean())
        assert rel_l2 <= max_l2 and frac <= max_frac, f"{what}: rel L2 {rel_l2:.2e}, {frac:.2e} of the elements off"

    # Outputs are compared in the max norm.  Gradients pass through the sampling locations: grad_sampling_loc is discontinuous
    # where a sample sits on a pixel-centre line.  Two implementations of the Linear layers differ in the
    # last bits (cuBLAS vs MKL: ~1e-7; 3xTF32 vs either: ~1e-6), so of the 7.7 M samples of this stack a handful land on the
    # other side of such a line and change the gradient of THEIR query row by a few percent (tools/stack_debug.py: 17 of 7650
    # rows with the tensor-core Linear layers, 0-1 with torch's).  A max-norm bound is the wrong metric for that: gradients are
    # held to a relative L2 error, plus a bound on the share of elements off by more than 1e-3 of the largest entry for the
    # per-query input gradients (one of the 100 clip queries is 1 %).  The gradient of the offset projections is the sum of
    # grad_offsets x query over ~10^2..10^4 queries and feels a single flip most.
    for tag, (outs, gin, gp), out_tol, l2, l2_off, frac in (("torch Linear", exact, 2e-5, 5e-4, 5e-3, 5e-3),
                                                            ("as shipped", shipped, 1e-4, 2e-3, 2e-2, 2e-2)):
        for name, a, b in zip(("memory", "q_box", "q_inst"), outs, outs_cpu):
            assert nerr(a, b) <= out_tol, f"{tag} {name}: {nerr(a, b):.2e}"
        for k in grad_names:
            close(gin[k], cpu[k].grad, f"{tag} grad {k}", l2, frac)
        for k in pg:
            close(gp[k], pg[k], f"{tag} grad {k}", l2_off if "offsets" in k else l2, 1.0)
