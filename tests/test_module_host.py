"""Host-side logic of the MSDeformAttn module (sampling-location arithmetic, softmax, temporal
level_start encoding, parameter initialisation, state-dict keys) checked on CPU: the module runs with
the C-oracle Function substituted for the CUDA one and must reproduce the fixtures recorded from the
unmodified reference module (tests/golden/make_golden.py section 5)."""
import pytest
import torch

import mdqe_cvpr2023_b200.modules as M
from tests.helpers import OracleMSDAFunction, load_golden, module_from_golden, module_inputs, nerr

CASES = ["module_spatial_pred", "module_spatial_grid", "module_temporal_grid"]


@pytest.fixture()
def oracle_function(monkeypatch):
    monkeypatch.setattr(M, "MSDeformAttnFunction", OracleMSDAFunction)


@pytest.mark.parametrize("name", CASES)
def test_module_matches_reference_module(name, oracle_function):
    z = load_golden(name)
    mod = module_from_golden(z, M.MSDeformAttn)
    query, ref, inp, shapes, mask = module_inputs(z)
    out = mod(query, ref, inp, shapes, mask)
    assert nerr(out, z["out"]) <= 2e-5
    out.backward(torch.from_numpy(z["grad_out"]))
    assert nerr(query.grad, z["grad_query"]) <= 5e-5
    assert nerr(inp.grad, z["grad_input"]) <= 5e-5
    for k, p in mod.named_parameters():
        assert nerr(p.grad, z["gp." + k]) <= 1e-4, k


def test_state_dict_keys_match_reference():
    z = load_golden("module_spatial_pred")
    enc = M.MSDeformAttn(d_model=64, n_levels=4, n_heads=4, n_points=4, pred_offsets=True, mode="spatial")
    assert sorted(enc.state_dict().keys()) == sorted(k[3:] for k in z.files if k.startswith("sd."))
    z = load_golden("module_temporal_grid")
    dec = M.MSDeformAttn(d_model=64, n_levels=4, n_heads=4, n_points=4, n_frames=3, pred_offsets=False, mode="temporal")
    assert sorted(dec.state_dict().keys()) == sorted(k[3:] for k in z.files if k.startswith("sd."))
    for k, v in dec.state_dict().items():
        assert tuple(v.shape) == tuple(z["sd." + k].shape), k


def test_reset_parameters_initial_pattern():
    # encoder: zero weight, bias = head ray * (k+1)/K * 8 * 0.05 * (l+1)   (ms_deform_attn.py:80-92)
    m = M.MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4, pred_offsets=True)
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, :, :, 0], (torch.arange(1, 5).view(1, 4) / 4 * 8 * 0.05 * torch.arange(1, 5).view(4, 1)).float())
    assert torch.allclose(b[0, :, :, 1], torch.zeros(4, 4), atol=1e-6)
    assert float(m.attention_weights.weight.abs().max()) == 0.0 and float(m.attention_weights.bias.abs().max()) == 0.0
    # decoder: buffer holds the unscaled ray grid, learned residual starts at zero
    d = M.MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4, n_frames=4, pred_offsets=False, mode="temporal")
    assert tuple(d.sampling_offsets.shape) == (1, 1, 8, 4, 4, 2)
    assert torch.allclose(d.sampling_offsets[0, 0, 2, 0, :, 1], torch.arange(1, 5).float() / 4 * 8)
    assert list(d.lvl_spatial_scales) == [2, 2, 2, 2]
    d._reset_parameters()     # the decoder calls it a second time (transformer_dec.py:72-74)


def test_constructor_validation():
    with pytest.raises(ValueError):
        M.MSDeformAttn(d_model=250, n_heads=8)
    with pytest.raises(ValueError):
        M.MSDeformAttn(mode="spatio-temporal")
