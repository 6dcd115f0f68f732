"""CPU model of the paired-corner value layout and of the GEMM epilogue that writes it (csrc/msda_common.cuh PackedLevel,
csrc/msda_packed.cu msda_pack_value_kernel, csrc/gemm3x.cuh packed epilogue).

Definition (the pack pass):  packed[n][prow][m] = { value[n, cell(y, xp - 1), m, :], value[n, cell(y, xp), m, :] },
prow = pstart_l + y (W_l + 1) + xp, xp in [0, W_l], cells outside the level row are zeros.

The value_proj epilogue produces the same tensor from the other side: the thread that owns pixel (n, s) of the GEMM output writes
its head slice into the RIGHT half of line xp = x and the LEFT half of line xp = x + 1 of its level row, plus the zero halves at the
two ends of the row.  This restates both addressings in numpy and checks that the scatter form writes every half line of the table
exactly once and builds exactly the gather form's tensor -- for ragged pyramids, 1 x 1 levels, a level whose table entry does not fit
(disabled: no lines), and rows that belong to no level.  (The CUDA code is checked on the GPU:
tests/test_msda_gpu.py::test_value_proj_epilogue_writes_the_packed_layout.)"""
import numpy as np
import pytest


def level_table(pyr, starts, S):
    """stage_packed_levels: (H, W, start, pstart) per level; a level that does not fit S rows (or the 2 S line budget) is disabled"""
    out, p = [], 0
    for (H, W), st in zip(pyr, starts):
        ok = H >= 0 and W >= 0 and st >= 0 and H * W <= S and st <= S - H * W and p + H * (W + 1) <= 2 * S
        out.append((H if ok else 0, W if ok else 0, st if ok else 0, p))
        p += (H * (W + 1)) if ok else 0
    return out, p


def pack_by_gather(value, pyr, starts):
    """msda_pack_value_kernel: one pass over the lines of the table"""
    N, S, M, D = value.shape
    table, total = level_table(pyr, starts, S)
    packed = np.zeros((N, 2 * S, M, 2, D), value.dtype)
    for n in range(N):
        for prow in range(total):
            lvl = 0
            while lvl + 1 < len(table) and prow >= table[lvl + 1][3]:
                lvl += 1
            H, W, st, ps = table[lvl]
            q = prow - ps
            y, xp = divmod(q, W + 1)
            for half, x in ((0, xp - 1), (1, xp)):
                if 0 <= x < W and y < H:
                    packed[n, prow, :, half] = value[n, st + y * W + x]
    return packed, total


def pack_by_scatter(value, pyr, starts):
    """the GEMM epilogue: one thread per pixel row (n, s); returns the tensor and how often each half line was written"""
    N, S, M, D = value.shape
    table, total = level_table(pyr, starts, S)
    packed = np.full((N, 2 * S, M, 2, D), np.nan, value.dtype)
    writes = np.zeros((N, 2 * S, 2), np.int64)
    for n in range(N):
        for s in range(S):
            line = None
            for H, W, st, ps in table:
                q = s - st
                if 0 <= q < H * W:
                    y, x = divmod(q, W)
                    line, first_x, last_x = ps + y * (W + 1) + x, x == 0, x == W - 1
                    break
            if line is None:
                continue                                    # a row outside every level writes nothing
            packed[n, line, :, 1] = value[n, s]             # right half of line xp = x
            packed[n, line + 1, :, 0] = value[n, s]         # left half of line xp = x + 1
            writes[n, line, 1] += 1
            writes[n, line + 1, 0] += 1
            if first_x:
                packed[n, line, :, 0] = 0                   # left half of line xp = 0
                writes[n, line, 0] += 1
            if last_x:
                packed[n, line + 1, :, 1] = 0               # right half of line xp = W
                writes[n, line + 1, 1] += 1
    return packed, writes, total


CASES = [
    ([(12, 20), (6, 10), (3, 5), (2, 3)], None, 0),
    ([(7, 9), (5, 4), (1, 1), (2, 6)], None, 0),
    ([(1, 1)], None, 0),
    ([(3, 1), (1, 4)], None, 0),
    ([(4, 5), (2, 3)], None, 3),                            # three rows at the end of S belong to no level
    ([(4, 5), (9, 9), (2, 3)], "second level does not fit", 0),
]


@pytest.mark.parametrize("pyr,broken,extra", CASES)
def test_scatter_form_builds_the_gather_form(pyr, broken, extra):
    rng = np.random.default_rng(len(pyr) * 17 + extra)
    sizes = [h * w for h, w in pyr]
    if broken:
        S = sizes[0] + sizes[2]                             # the table claims 81 rows for level 1 that the tensor does not have
        starts = [0, sizes[0], sizes[0]]
    else:
        S = sum(sizes) + extra
        starts = list(np.cumsum([0] + sizes[:-1]))
    N, M, D = 2, 3, 4
    value = rng.standard_normal((N, S, M, D)).astype(np.float32)
    want, total = pack_by_gather(value, pyr, starts)
    got, writes, total2 = pack_by_scatter(value, pyr, starts)
    assert total == total2 <= 2 * S
    assert (writes[:, :total] == 1).all(), "every half line of the table is written exactly once"
    assert (writes[:, total:] == 0).all(), "nothing is written past the table"
    np.testing.assert_array_equal(got[:, :total], want[:, :total])
    if broken:
        table, _ = level_table(pyr, starts, S)
        assert table[1][:2] == (0, 0), "the level that does not fit is disabled, not dereferenced"


def test_sampler_reads_two_lines_per_sample():
    """what the layout is for: the four corners (y0 | y0 + 1) x (x0 | x0 + 1) of a bilinear sample are the two halves of the lines
    (y0, xp = x0 + 1) and (y0 + 1, xp = x0 + 1), and a corner outside the row reads the zeros stored there"""
    pyr, starts = [(5, 7)], [0]
    rng = np.random.default_rng(0)
    value = rng.standard_normal((1, 35, 1, 2)).astype(np.float32)
    packed, _ = pack_by_gather(value, pyr, starts)
    H, W = pyr[0]
    for y0 in range(-1, H):
        for x0 in range(-1, W):
            for dy in (0, 1):
                y = y0 + dy
                if not 0 <= y < H:
                    continue
                line = y * (W + 1) + x0 + 1
                for dx in (0, 1):
                    x = x0 + dx
                    want = value[0, y * W + x, 0] if 0 <= x < W else np.zeros(2, np.float32)
                    np.testing.assert_array_equal(packed[0, line, 0, dx], want)
