"""CPU model of the backward's merged grad_value reductions (csrc/msda_fast2.cuh, merge_level_slots): the lanes of one
(pair, level) exchange a packed cell coordinate and their four corner weights; a lane adds the partner's 2x2 weights shifted by
the cell difference and drops every corner a lower lane also holds.  This restates that lane program in numpy, statement by
statement, and checks it against the definition it implements: group the 4*P corner records by value row, the lowest lane keeps
the row with the summed weight, everybody else gets 0 -- including out-of-range corners, the "insane sample" marker (-8, -8) and
lanes without a sample.  (The CUDA code itself is checked on the GPU: tests/test_msda_gpu.py::test_backward_merged_reductions.)"""
import itertools

import numpy as np
import pytest

NO_MERGE_KEY = 0xFFFF8000
INVALID = 0xFFFFFFFF


def lane_program(keys, w, P):
    """keys[P] uint32, w[P,4] float64 (corner = dy*2+dx) -> merged weights [P,4] as the kernel computes them."""
    out = np.zeros_like(w)
    for me in range(P):
        cx, cy = int(keys[me] & 0xFFFF), int(keys[me] >> 16)
        wsum = w[me].copy()
        kill = 0
        for j in range(1, P):
            partner = me ^ j
            pk, pw = int(keys[partner]), w[partner]
            ex, ey = (pk & 0xFFFF) - cx, (pk >> 16) - cy
            x0, xm, xp, y0, ym, yp = ex == 0, ex == -1, ex == 1, ey == 0, ey == -1, ey == 1
            a00 = pw[0] if x0 else (pw[1] if xm else 0.0)
            a01 = pw[1] if x0 else (pw[0] if xp else 0.0)
            a10 = pw[2] if x0 else (pw[3] if xm else 0.0)
            a11 = pw[3] if x0 else (pw[2] if xp else 0.0)
            wsum[0] += a00 if y0 else (a10 if ym else 0.0)
            wsum[1] += a01 if y0 else (a11 if ym else 0.0)
            wsum[2] += a10 if y0 else (a00 if yp else 0.0)
            wsum[3] += a11 if y0 else (a01 if yp else 0.0)
            if partner < me:
                mx0, mx1, my0, my1 = x0 or xm, x0 or xp, y0 or ym, y0 or yp
                kill |= (1 if mx0 and my0 else 0) | (2 if mx1 and my0 else 0) | (4 if mx0 and my1 else 0) | (8 if mx1 and my1 else 0)
        for c in range(4):
            out[me, c] = 0.0 if (kill >> c) & 1 else wsum[c]
    return out


def definition(cells, valid, w, P):
    """cells[P,2] (x0, y0), valid[P,4], w[P,4]: lowest lane holding a (valid) row keeps the sum."""
    out = np.zeros_like(w)
    owner = {}
    for lane, c in itertools.product(range(P), range(4)):
        if not valid[lane, c]:
            continue
        cell = (cells[lane, 0] + (c & 1), cells[lane, 1] + (c >> 1))
        if cell not in owner:
            owner[cell] = (lane, c)
        out[owner[cell]] += w[lane, c]
    return out


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("H,W", [(1, 1), (2, 3), (6, 10), (48, 80)])
def test_lane_program_equals_row_grouping(P, H, W):
    rng = np.random.default_rng(1000 * P + 10 * H + W)
    for trial in range(400):
        spread = rng.choice([0, 1, 2, max(H, W)])
        base = np.array([rng.integers(-1, W), rng.integers(-1, H)])
        cells = base + rng.integers(-spread, spread + 1, size=(P, 2))
        cells[:, 0] = np.clip(cells[:, 0], -1, W - 1)
        cells[:, 1] = np.clip(cells[:, 1], -1, H - 1)
        has_sample = rng.random(P) < 0.9
        insane = rng.random(P) < 0.05                      # NaN / far-away location: the kernel marks the cell (-8, -8)
        cells[insane] = -8
        valid = np.zeros((P, 4), bool)
        for lane, c in itertools.product(range(P), range(4)):
            x, y = cells[lane, 0] + (c & 1), cells[lane, 1] + (c >> 1)
            valid[lane, c] = has_sample[lane] and 0 <= x < W and 0 <= y < H
        w = np.where(valid, rng.integers(1, 64, size=(P, 4)).astype(np.float64), 0.0)     # integers: sums are exact in any order
        keys = np.array([((int(cells[l, 0]) + 8) | ((int(cells[l, 1]) + 8) << 16)) if has_sample[l] else NO_MERGE_KEY + 4 * l
                         for l in range(P)], dtype=np.uint64)
        got = lane_program(keys, w, P)
        want = definition(cells, valid, w, P)
        assert np.array_equal(got, want), (trial, cells, valid, w, got, want)
        assert got.sum() == w.sum()                        # nothing lost, nothing counted twice


def test_oversized_levels_and_missing_samples_never_match():
    # lanes of a level wider than the 15-bit key fields (or without a sample) carry distinct no-merge keys: nothing merges
    P = 4
    keys = np.array([NO_MERGE_KEY + 4 * l for l in range(P)], dtype=np.uint64)
    w = np.arange(1, 17, dtype=np.float64).reshape(P, 4)
    assert np.array_equal(lane_program(keys, w, P), w)
    # a real cell next to the largest representable coordinate does not alias a no-merge key
    keys[0] = (32751 + 8) | ((32751 + 8) << 16)
    assert np.array_equal(lane_program(keys, w, P), w)
