"""CPU: the rasteriser's work items (emit_items in csrc/smalfit_kernels.cu, the draw loop of
csrc/smalfit_raster_tile.cuh), restated in numpy.  This does not execute the CUDA code (the GPU parity tests do); it
pins what the kernels rely on:

  * every non-empty tile yields 1, 2, 4 or 8 band items that cover its 32 rows exactly once,
  * an item never exceeds the size limit unless its tile is already cut into 8 bands,
  * a size class can hold every item of a launch (frames x tiles x 8 slots), for any arrival order,
  * walking the classes end to end with one counter hands out every item exactly once, larger classes first,
  * the item size rule follows the PREVIOUS launch's pair total, clamped to [RT_MIN_ITEM, RT_MAX_ITEM].
"""
import numpy as np
import pytest

RT_ITEM_BINS, RT_FAIR, RT_MIN_ITEM, RT_MAX_ITEM, TILE_H = 24, 2, 16000, 64000, 32


def item_size_limit(prev_total, n_ctas, fair=0, min_item=0, split_len=0):
    share = prev_total // ((fair if fair > 0 else RT_FAIR) * n_ctas)
    lo = min_item if min_item > 0 else RT_MIN_ITEM
    hi = (1 << 30) if fair > 0 else RT_MAX_ITEM
    return split_len if split_len > 0 else min(max(share, lo), hi)


def emit_items(cost, cmax, list_cap=192 * 1024):
    """(log2 of the band count, size class) of one tile."""
    lg = 0
    while lg < 3 and ((cost >> lg) > cmax or (cost >> lg) > list_cap):
        lg += 1
    r = np.float32(cost >> lg) / np.float32(cmax)
    if r >= 1:
        return lg, 0
    return lg, min(1 + int(np.float32(-2.0) * np.log2(max(r, np.float32(1e-6)))), RT_ITEM_BINS - 1)


@pytest.mark.parametrize("n_frames,tiles,seed", [(16, 64, 0), (128, 64, 1), (3, 256, 2), (1, 1024, 3)])
def test_items_cover_every_tile_once_and_are_handed_out_once(n_frames, tiles, seed):
    rng = np.random.default_rng(seed)
    cost = np.where(rng.random((n_frames, tiles)) < 0.2, rng.integers(1, 400000, size=(n_frames, tiles)), 0)
    cost[0, 0] = 7_900_000                                   # a tile every face reaches
    n_ctas = 444
    cmax = item_size_limit(int(cost.sum()), n_ctas)
    assert RT_MIN_ITEM <= cmax <= RT_MAX_ITEM
    bin_cap = n_frames * tiles * 8
    bins = [[] for _ in range(RT_ITEM_BINS)]
    order = rng.permutation(n_frames)                          # frames finish in any order
    for f in order:
        for t in range(tiles):
            c = int(cost[f, t])
            if c == 0:
                continue                                       # frame_front finished it
            lg, b = emit_items(c, cmax)
            assert 0 <= b < RT_ITEM_BINS
            per = c >> lg
            assert per <= cmax or lg == 3
            for band in range(1 << lg):
                bins[b].append((f, t, band, lg, per))
    assert all(len(b) <= bin_cap for b in bins)
    # draw loop: k-th draw = k-th item of the classes laid end to end
    ends = np.cumsum([len(b) for b in bins])
    rows = np.zeros((n_frames, tiles, TILE_H), np.int32)
    last_class = 0
    for k in range(int(ends[-1])):
        b = int(np.searchsorted(ends, k, side="right"))
        f, t, band, lg, per = bins[b][k - (int(ends[b - 1]) if b else 0)]
        assert b >= last_class
        last_class = b
        bh = TILE_H >> lg
        rows[f, t, band * bh:(band + 1) * bh] += 1
    assert np.array_equal(rows.max(axis=2), (cost > 0).astype(np.int32))
    assert np.array_equal(rows.min(axis=2), (cost > 0).astype(np.int32))
    # classes are ordered by size: nothing in a later class is larger than the smallest item two classes earlier
    sizes = [[it[4] for it in b] for b in bins]
    for b in range(2, RT_ITEM_BINS):
        if sizes[b] and sizes[b - 2]:
            assert max(sizes[b]) <= min(sizes[b - 2])


def test_item_size_rule():
    assert item_size_limit(0, 444) == RT_MIN_ITEM                        # first launch: nothing known yet
    assert item_size_limit(13_600_000, 444) == RT_MIN_ITEM               # 16 frames per GPU: floor
    assert item_size_limit(108_600_000, 444) == RT_MAX_ITEM              # 128 frames: cap (lists of the items in flight ~ L2)
    assert item_size_limit(40_000_000, 444) == 40_000_000 // 888
    assert item_size_limit(108_600_000, 444, fair=1) == 108_600_000 // 444   # an explicit fair share is taken literally
    assert item_size_limit(108_600_000, 444, split_len=64) == 64
