"""CPU model of how the Linear-layer GEMM deals its work to CTAs (csrc/gemm3x.cuh: G3Walk, g3_tile_parts).

The kernel cuts the (tile, chunk) units of a GEMM into one contiguous range per CTA ("stream-K").  A tile whose chunks fall into
several ranges is combined in the output: the stretch holding the tile's LAST chunks stores (and carries the bias) and publishes a
flag, every other stretch waits for the flag and reduce-adds; the arrival that sees `2 * parts - 1` in the flag resets it.  This
restates the walk statement by statement and checks what the protocol relies on:

  * every unit is computed exactly once and a stretch never crosses a tile,
  * a tile has exactly one storing stretch, and it is the first stretch of its CTA unless it is a whole tile
    (so nobody waits for work that is scheduled after its own: the wait cannot deadlock),
  * `g3_tile_parts` equals the number of stretches of the tile (else a flag would never be reset, or reset early),
  * the round-robin forms (whole tiles; the weight gradient's split reduction added into a zeroed output) partition the work too,
    and the weight gradient never gives an SM two items.

(The CUDA code is checked on the GPU: tests/test_linear_gpu.py runs both distributions against fp64.)"""
import pytest

STORE, REDUCE, STORE_PUBLISH, WAIT_REDUCE = 0, 1, 2, 3
MODE_STORE, MODE_REDUCE, MODE_STREAM_K = 0, 1, 2


def walk(mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid, cta):
    """the stretches (tm, tn, c0, c1, out, tile, with_bias) CTA `cta` of `grid` computes, in order (G3Walk)"""
    out = []
    if mode == MODE_STREAM_K:
        q, r = n_items // grid, n_items % grid
        cur = cta * q + min(cta, r)
        end = cur + q + (1 if cta < r else 0)
        while cur < end:
            tile = cur // n_kchunks
            c0 = cur - tile * n_kchunks
            c1 = min(n_kchunks, c0 + (end - cur))
            cur += c1 - c0
            tm = tile // tiles_n
            tn = tile - tm * tiles_n
            kind = (STORE if c0 == 0 else STORE_PUBLISH) if c1 == n_kchunks else WAIT_REDUCE
            out.append((tm, tn, c0, c1, kind, tile, kind != WAIT_REDUCE))
    else:
        cur = cta
        while cur < n_items:
            r = cur // tiles_n
            split = r // tiles_m
            tn, tm = cur - r * tiles_n, r - split * tiles_m
            c0 = split * cps
            out.append((tm, tn, c0, min(n_kchunks, c0 + cps), mode, 0, split == 0))
            cur += grid
    return out


def tile_parts(tile, n_kchunks, n_units, grid):
    q, r = n_units // grid, n_units % grid
    cta_of = lambda u: u // (q + 1) if u < r * (q + 1) else r + (u - r * (q + 1)) // q
    return cta_of(tile * n_kchunks + n_kchunks - 1) - cta_of(tile * n_kchunks) + 1


def host_plan(M, N, K, splits, sms=148, stream_k=True):
    """launch_gemm3x (csrc/mask_gemm.cu): -> (mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid)"""
    n_kchunks = (K + 31) // 32
    tiles_m, tiles_n = (M + 127) // 128, (N + 127) // 128
    splits = max(1, min(splits, n_kchunks))
    cps = (n_kchunks + splits - 1) // splits
    splits = (n_kchunks + cps - 1) // cps
    tiles = tiles_m * tiles_n
    n_items, mode = tiles * splits, (MODE_REDUCE if splits > 1 else MODE_STORE)
    idle_units = ((tiles + sms - 1) // sms * sms - tiles) * n_kchunks
    if splits == 1 and stream_k and (tiles < sms or idle_units >= 4 * sms or stream_k == 2):
        mode, n_items = MODE_STREAM_K, tiles * n_kchunks
    return mode, n_kchunks, cps, tiles_m, tiles_n, n_items, min(n_items, sms)


CASES = [  # (M, N, K, splits): the module's GEMMs at the BASELINE shapes and the odd shapes of tests/test_linear_gpu.py
    (20400, 256, 256, 1), (20400, 384, 256, 1), (20400, 128, 256, 1), (20400, 256, 384, 1), (61200, 256, 256, 1), (15300, 192, 192, 1),
    (784, 256, 256, 1), (784, 384, 256, 1), (196, 256, 256, 1), (1, 256, 256, 1), (130, 20, 36, 1), (257, 300, 8, 1), (1000, 4, 260, 1),
    (148 * 128, 128, 64, 1), (300, 96, 512, 1),
    (256, 256, 20400, 37), (384, 256, 20400, 24), (128, 256, 20400, 74), (256, 256, 784, 37), (20, 36, 130, 148), (300, 8, 257, 49), (4, 260, 1000, 49),
]


@pytest.mark.parametrize("stream_k", [1, 2, 0])
@pytest.mark.parametrize("M,N,K,splits", CASES)
def test_every_unit_once_and_one_storer_per_tile(M, N, K, splits, stream_k):
    mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid = host_plan(M, N, K, splits, stream_k=stream_k)
    assert 1 <= grid <= 148
    seen = {}
    stretches_of_tile = {}
    for cta in range(grid):
        segs = walk(mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid, cta)
        if mode == MODE_STREAM_K:
            assert segs, "every CTA of the contiguous form has work"
        for i, (tm, tn, c0, c1, kind, tile, with_bias) in enumerate(segs):
            assert 0 <= tm < tiles_m and 0 <= tn < tiles_n and 0 <= c0 < c1 <= n_kchunks
            for c in range(c0, c1):
                assert (tm, tn, c) not in seen, "a unit computed twice"
                seen[(tm, tn, c)] = cta
            stretches_of_tile.setdefault((tm, tn), []).append((cta, i, c0, c1, kind, with_bias))
            if mode == MODE_STREAM_K:
                assert tile == tm * tiles_n + tn
                if kind == STORE_PUBLISH:
                    assert i == 0, "the publishing stretch must be its CTA's first work"
                if kind == WAIT_REDUCE:
                    assert i == len(segs) - 1, "a waiting stretch is its CTA's last work"
    assert len(seen) == tiles_m * tiles_n * n_kchunks, "a unit nobody computes"
    for (tm, tn), st in stretches_of_tile.items():
        kinds = [k for _, _, _, _, k, _ in st]
        assert sum(1 for *_, b in st if b) == 1, "the bias is added exactly once per tile"
        if mode == MODE_STREAM_K:
            assert sum(1 for k in kinds if k in (STORE, STORE_PUBLISH)) == 1
            if len(st) > 1:
                assert STORE not in kinds and kinds.count(STORE_PUBLISH) == 1 and kinds.count(WAIT_REDUCE) == len(st) - 1
                publisher = next(c for c, _, _, _, k, _ in st if k == STORE_PUBLISH)
                assert all(c < publisher for c, _, _, _, k, _ in st if k == WAIT_REDUCE), "waiters hold the earlier chunks, in lower CTAs"
            assert tile_parts(tm * tiles_n + tn, n_kchunks, n_items, grid) == len(st)
        elif mode == MODE_STORE:
            assert kinds == [STORE]
        else:
            assert all(k == REDUCE for k in kinds)


def test_weight_gradient_items_fit_one_round():
    for out_f, in_f in ((256, 256), (384, 256), (128, 256), (192, 192), (288, 192), (96, 192)):
        tiles = ((out_f + 127) // 128) * ((in_f + 127) // 128)
        mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid = host_plan(out_f, in_f, 20400, max(2, 148 // tiles))
        assert mode == MODE_REDUCE and n_items <= 148 and n_items > 148 - tiles - 8, (out_f, in_f, n_items)


def test_balance_of_the_contiguous_form():
    """what the form is for: 320 tiles on 148 SMs are 24 chunks for the busiest CTA round-robin, 18 in contiguous ranges"""
    mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid = host_plan(20400, 256, 256, 1)
    load = [sum(c1 - c0 for _, _, c0, c1, *_ in walk(mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid, c)) for c in range(grid)]
    assert max(load) == 18 and min(load) == 17
    assert host_plan(15300, 192, 192, 1)[0] == MODE_STORE and host_plan(15300, 192, 192, 1, stream_k=2)[0] == MODE_STREAM_K
    mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid = host_plan(20400, 256, 256, 1, stream_k=0)
    load = [sum(c1 - c0 for _, _, c0, c1, *_ in walk(mode, n_kchunks, cps, tiles_m, tiles_n, n_items, grid, c)) for c in range(grid)]
    assert max(load) == 24 and min(load) == 16
