"""Pin the plain-C oracle (oracle/msda_oracle.c) against fixtures produced by the reference's own
Python (tests/golden/make_golden.py: ms_deform_attn_core_pytorch + autograd, and the mask einsum).

Tolerances: fp64 fixtures 1e-11 normalised max error (summation order only); fp32 fixtures 2e-6.
The reference's own fixture tolerance (ops/test.py:40, torch.allclose rtol 1e-5 / atol 1e-8 in
double; :56 rtol 1e-2 / atol 1e-3 in float) is checked as well.
"""
import glob
import os

import numpy as np
import pytest

from oracle import msda_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CORE = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
              if not os.path.basename(p).startswith(("module_", "mask_", "mask_losses_", "match_cost_", "nms_siou_", "track_siou_", "aligned_bilinear_", "query_init_")))


def nerr(a, b):
    denom = np.abs(b).max()
    return np.abs(a - b).max() / (denom if denom > 0 else 1.0)


def test_fixture_inventory():
    assert "ref_fixture_D2_f64" in CORE and "pyramid_D32_f32" in CORE and "edge_coords_f64" in CORE
    assert len(CORE) >= 12


@pytest.mark.parametrize("name", CORE)
def test_core_against_reference_python(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    tol = 1e-11 if z["value"].dtype == np.float64 else 2e-6
    out = O.msda_forward(z["value"], z["shapes"], z["loc"], z["aw"], z["level_start"])
    assert out.shape == z["out"].shape
    assert nerr(out, z["out"]) <= tol
    if name.startswith("ref_fixture"):
        if z["value"].dtype == np.float64:
            assert np.allclose(out, z["out"], rtol=1e-5, atol=1e-8)      # ops/test.py:40
        else:
            assert np.allclose(out, z["out"], rtol=1e-2, atol=1e-3)      # ops/test.py:56
    gv, gl, ga = O.msda_backward(z["value"], z["shapes"], z["loc"], z["aw"], z["grad_out"], z["level_start"])
    assert nerr(gv, z["grad_value"]) <= tol
    assert nerr(gl, z["grad_loc"]) <= tol
    assert nerr(ga, z["grad_aw"]) <= tol


def test_edge_coords_have_zero_and_nonzero_samples():
    z = np.load(os.path.join(GOLDEN, "edge_coords_f64.npz"))
    out = O.msda_forward(z["value"], z["shapes"], z["loc"], z["aw"])
    rows = np.abs(out[0]).sum(-1)
    assert (rows == 0).any() and (rows > 0).any()     # far-outside samples contribute exactly 0


@pytest.mark.parametrize("name", ["mask_einsum_K32", "mask_einsum_K24"])
def test_mask_against_einsum(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = O.mask_forward(z["coeff"], z["proto"])
    assert nerr(out, z["out"]) <= 2e-6
    gc, gp = O.mask_backward(z["coeff"], z["proto"], z["grad_out"])
    assert nerr(gc, z["grad_coeff"]) <= 2e-6
    assert nerr(gp, z["grad_proto"]) <= 2e-6


def test_linearity_in_value_and_weights():
    rng = np.random.default_rng(0)
    shapes = np.array([[5, 7], [3, 4]], dtype=np.int64)
    S = 35 + 12
    v1 = rng.standard_normal((2, S, 3, 8)); v2 = rng.standard_normal((2, S, 3, 8))
    loc = rng.random((2, 6, 3, 2, 3, 2)) * 1.2 - 0.1
    aw = rng.random((2, 6, 3, 2, 3))
    a = O.msda_forward(v1, shapes, loc, aw); b = O.msda_forward(v2, shapes, loc, aw)
    c = O.msda_forward(2 * v1 - 3 * v2, shapes, loc, aw)
    assert nerr(c, 2 * a - 3 * b) < 1e-12
    assert nerr(O.msda_forward(v1, shapes, loc, 0.5 * aw), 0.5 * a) < 1e-12


def test_empty_query_set():
    shapes = np.array([[2, 2]], dtype=np.int64)
    out = O.msda_forward(np.ones((1, 4, 1, 4)), shapes, np.zeros((1, 0, 1, 1, 1, 2)), np.zeros((1, 0, 1, 1, 1)))
    assert out.shape == (1, 0, 4)


@pytest.mark.parametrize("name", [n for n in CORE if n.endswith("f32")] + ["edge_coords_f64"])
def test_torch_port_against_reference_python(name):
    """oracle/torch_port.py (the CPU baseline bench.py times) reproduces the reference's function."""
    import torch
    from oracle import torch_port as TP
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = {k: torch.from_numpy(z[k]) for k in ("value", "shapes", "level_start", "loc", "aw", "grad_out")}
    out, gv, gl, ga = TP.msda_fwd_bwd_torch(t["value"], t["shapes"], t["loc"], t["aw"], t["grad_out"], t["level_start"])
    tol = 1e-12 if z["value"].dtype == np.float64 else 1e-6
    for got, key in ((out, "out"), (gv, "grad_value"), (gl, "grad_loc"), (ga, "grad_aw")):
        assert nerr(got.numpy(), z[key]) <= tol, key
