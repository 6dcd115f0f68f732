"""Drop-in level L2 inside the reference's OWN caller code, on CPU in the build container:
the unmodified reference `Transformer_Enc` (transformer_enc.py) and `DecoderDefAttnLayer` (transformer_dec.py) are built
twice -- once with the reference's MSDeformAttn, once with this package's -- with identical weights (state dicts must
load across), and must produce the same outputs and gradients.  Both run on the C oracle Function (no GPU here); the
CUDA side of the same module is covered by tests/test_dropin_gpu.py.  Skipped where /root/reference does not exist."""
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mdqe")), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def ref_modules():
    saved = {k: sys.modules.get(k) for k in ("MultiScaleDeformableAttention", "mdqe", "mdqe.models", "mdqe.models.ops",
                                             "mdqe.util")}
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    for name, sub in (("mdqe", "mdqe"), ("mdqe.models", "mdqe/models"), ("mdqe.models.ops", "mdqe/models/ops"),
                      ("mdqe.util", "mdqe/util")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = mod
    import mdqe.models.ops.modules.ms_deform_attn as ref_attn
    import mdqe.models.transformer_dec as ref_dec
    import mdqe.models.transformer_enc as ref_enc
    from tests.helpers import OracleMSDAFunction
    ref_attn.MSDeformAttnFunction = OracleMSDAFunction
    yield ref_attn, ref_enc, ref_dec
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for k in [k for k in sys.modules if k.startswith("mdqe.")]:
        sys.modules.pop(k, None)


@pytest.fixture()
def ours(monkeypatch):
    import mdqe_cvpr2023_b200.modules as M
    from tests.helpers import OracleMSDAFunction
    monkeypatch.setattr(M, "MSDeformAttnFunction", OracleMSDAFunction)
    return M


def _pyramid(B, dim, g):
    shapes = [(8, 12), (4, 6), (2, 3), (1, 2)]
    srcs = [torch.randn(B, dim, h, w, generator=g) for h, w in shapes]
    masks = [torch.rand(B, h, w, generator=g) < 0.1 for h, w in shapes]
    pos = [torch.randn(B, dim, h, w, generator=g) for h, w in shapes]
    return shapes, srcs, masks, pos


def test_reference_encoder_runs_unchanged_on_our_module(ref_modules, ours):
    ref_attn, ref_enc, _ = ref_modules
    torch.manual_seed(0)
    enc_ref = ref_enc.Transformer_Enc(64, n_heads=4, n_feature_levels=4, n_enc_points=4, n_enc_layers=2)
    ref_enc.MSAttnBlock = ours.MSDeformAttn                      # the one-line swap of INTEGRATION.md section 3
    try:
        enc_ours = ref_enc.Transformer_Enc(64, n_heads=4, n_feature_levels=4, n_enc_points=4, n_enc_layers=2)
    finally:
        ref_enc.MSAttnBlock = ref_attn.MSDeformAttn
    assert isinstance(enc_ours.encoder.layers[0].self_attn, ours.MSDeformAttn)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in enc_ref.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    enc_ours.load_state_dict(enc_ref.state_dict(), strict=True)   # checkpoints move across unchanged
    _, srcs, masks, pos = _pyramid(2, 64, g)
    srcs_a = [s.clone().requires_grad_(True) for s in srcs]
    srcs_b = [s.clone().requires_grad_(True) for s in srcs]
    out_ref = enc_ref(srcs_a, masks, pos)
    out_ours = enc_ours(srcs_b, masks, pos)
    assert torch.allclose(out_ref, out_ours, atol=2e-5, rtol=1e-4)
    w = torch.randn(out_ref.shape, generator=g)
    (out_ref * w).sum().backward()
    (out_ours * w).sum().backward()
    for a, b in zip(srcs_a, srcs_b):
        assert torch.allclose(a.grad, b.grad, atol=5e-5, rtol=1e-3)
    for (k, p), (_, q) in zip(enc_ref.named_parameters(), enc_ours.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-4, rtol=1e-3), k


def test_reference_decoder_layer_cross_attention_on_our_module(ref_modules, ours):
    """frame-level (spatial, box-scaled grid) and clip-level (temporal) cross attention of DecoderDefAttnLayer."""
    ref_attn, _, ref_dec = ref_modules
    T = 3
    torch.manual_seed(0)
    lay_ref = ref_dec.DecoderDefAttnLayer(64, 4, fpn_levels=4, n_frames=T, n_points=4, pred_offsets=False, use_tca=True)
    ref_dec.MSAttnBlock = ours.MSDeformAttn
    try:
        lay_ours = ref_dec.DecoderDefAttnLayer(64, 4, fpn_levels=4, n_frames=T, n_points=4, pred_offsets=False, use_tca=True)
    finally:
        ref_dec.MSAttnBlock = ref_attn.MSDeformAttn
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for p in lay_ref.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    lay_ours.load_state_dict(lay_ref.state_dict(), strict=True)
    lay_ref.train(); lay_ours.train()
    shapes_list = [(8, 12), (4, 6), (2, 3), (1, 2)]
    shapes = torch.tensor(shapes_list)
    S = sum(h * w for h, w in shapes_list)
    B, Q = 2, 9
    feats = torch.randn(B * T, S, 64, generator=g)
    pad = torch.rand(B * T, S, generator=g) < 0.1
    x = torch.randn(B * T, Q, 64, generator=g)
    xpos = torch.randn(B * T, Q, 64, generator=g)
    boxes = torch.cat([torch.rand(B * T, Q, 2, generator=g), torch.rand(B * T, Q, 2, generator=g) * 0.3 + 0.05], -1)
    a = lay_ref.forward_ca_box(x, xpos, boxes, feats, shapes, pad)
    b = lay_ours.forward_ca_box(x, xpos, boxes, feats, shapes, pad)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-4)
    inst_boxes = boxes[:B]
    inst_pos = torch.randn(B, Q, 64, generator=g)
    c = lay_ref.forward_ca_inst(x, a, None, inst_pos, inst_boxes, feats, shapes, pad)
    d = lay_ours.forward_ca_inst(x, b, None, inst_pos, inst_boxes, feats, shapes, pad)
    assert torch.allclose(c, d, atol=2e-5, rtol=1e-4)
