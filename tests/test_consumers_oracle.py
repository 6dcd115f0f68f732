"""The numpy oracle of the callers either side of the path (oracle/consumers_oracle.py) against fixtures produced by the
reference's own Python (tests/golden/make_golden_consumers.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import consumers_oracle as co

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def nerr(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", ["match_cost_K32", "match_cost_K24"])
def test_match_cost_oracle(name):
    d = load(name)
    bce, dice = co.match_cost(d["coeff"], d["proto"], d["targets"])
    assert nerr(bce, d["cost_bce"]) < 5e-6 and nerr(dice, d["cost_dice"]) < 5e-6      # fixtures are fp32


@pytest.mark.parametrize("name", ["nms_siou_T4", "nms_siou_T5"])
def test_nms_siou_oracle(name):
    d = load(name)
    assert nerr(co.nms_siou(d["mask_pred"]), d["siou"]) < 5e-6


@pytest.mark.parametrize("name", ["mask_losses_interinst", "mask_losses_plain"])
def test_mask_losses_oracle(name):
    d = load(name)
    ti = d["targets_interinst"] if int(d["interinst"]) else None
    lm, ld, gc, gp = co.mask_losses(d["coeff"][d["src_idx"]], d["proto"], d["targets"], ti, float(d["num_masks"]), d["grad_weights"])
    assert abs(lm - float(d["loss_mask"])) < 5e-6 * abs(float(d["loss_mask"])) and abs(ld - float(d["loss_dice"])) < 5e-6 * abs(float(d["loss_dice"]))
    want_gc = d["grad_coeff"][d["src_idx"]]
    assert nerr(gc, want_gc) < 1e-5 and nerr(gp, d["grad_proto"]) < 1e-5
    unmatched = np.setdiff1d(np.arange(d["coeff"].shape[0]), d["src_idx"])
    assert np.abs(d["grad_coeff"][unmatched]).max() == 0.0          # the reference's gradient lives on the matched rows only


@pytest.mark.parametrize("name", ["track_siou_a", "track_siou_b"])
def test_track_siou_oracle(name):
    d = load(name)
    assert nerr(co.track_siou(d["saved_masks"], d["input_masks"]), d["siou"]) < 5e-6


@pytest.mark.parametrize("name", ["aligned_bilinear_f4", "aligned_bilinear_f2"])
def test_aligned_bilinear_oracle(name):
    d = load(name)
    up = co.aligned_bilinear(d["x"], int(d["factor"]))
    assert up.shape == d["up"].shape
    assert nerr(up, d["up"]) < 2e-6
    assert nerr(1 / (1 + np.exp(-up)), d["up_sigmoid"]) < 2e-6


def test_query_init_oracle():
    d = load("query_init_f64")
    out, gf, gc = co.query_init_sample(d["feat"], d["shapes"], d["level_start"], d["coords"], d["grad_out"])
    assert nerr(out, d["out"]) < 1e-12
    assert nerr(gf, d["grad_feat"]) < 1e-12
    assert nerr(gc, d["grad_coords"]) < 1e-11
