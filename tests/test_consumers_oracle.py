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
