"""CPU: the tile rasteriser's *algorithm* (csrc/smalfit_raster_tile.cuh), restated in numpy and run against the
oracle.  This does not execute the CUDA code (the GPU parity tests do); it pins the design decisions the kernel
relies on, on a scene small enough for pure Python:

  * candidates per pixel from the integrated corner grid of the faces' rectangles == brute-force rectangle counts,
  * per-warp write cursors (list offset + candidates of the warps before) give every (face, pixel) pair its own slot,
    in face order, with no gaps other than the slots of rejected pairs,
  * K nearest by (depth, slot) == K nearest by (depth, face id) (the tile lists are in face order),
  * the result does not depend on how the list is split over warps, on multi-pass windows or on row bands,
  * and equals the oracle's soft silhouette.
"""
import math

import numpy as np
import pytest
import torch

from oracle import smal_oracle as O

TILE, WARPS = 32, 8


def _scene(seed, n_faces, S):
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-0.3, 0.3, size=(n_faces, 1, 2)) * np.array([1.0, 0.7])
    tri = ctr + rng.normal(size=(n_faces, 3, 2)) * 0.12
    z = rng.uniform(1.8, 3.2, size=(n_faces, 3, 1))
    xyz = np.concatenate([tri, z], axis=2)
    xyz[1::5] = xyz[0::5][: len(xyz[1::5])]            # exact copies of other faces: equal depths, ties at the K cut
    verts = xyz.reshape(-1, 3)
    faces = np.arange(n_faces * 3).reshape(n_faces, 3)
    return verts, faces


def _rects(verts, faces, S):
    """Conservative pixel rectangle per face (face_pixel_rect), None when empty."""
    r = math.sqrt(O.BLUR_RADIUS)
    out = []
    for f in faces:
        x, y = verts[f, 0], verts[f, 1]
        cl, ch = (1 - (x.max() + r)) * S / 2 - 0.5, (1 - (x.min() - r)) * S / 2 - 0.5
        rl, rh = (1 - (y.max() + r)) * S / 2 - 0.5, (1 - (y.min() - r)) * S / 2 - 0.5
        c0, c1 = int(max(math.ceil(cl - 0.01), 0)), int(min(math.floor(ch + 0.01), S - 1))
        r0, r1 = int(max(math.ceil(rl - 0.01), 0)), int(min(math.floor(rh + 0.01), S - 1))
        out.append((c0, c1, r0, r1) if c0 <= c1 and r0 <= r1 else None)
    return out


def _fragment(verts, face, px, py):
    """(depth, 1 - p) of one (face, pixel) pair or None (CheckPixelInsideFace, fp64)."""
    (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = verts[face]
    area = (x2 - x0) * (y1 - y0) - (y2 - y0) * (x1 - x0)
    if max(z0, z1, z2) < 0 or abs(area) <= O.K_EPS:
        return None
    den = area + O.K_EPS
    w0 = ((px - x1) * (y2 - y1) - (py - y1) * (x2 - x1)) / den
    w1 = ((px - x2) * (y0 - y2) - (py - y2) * (x0 - x2)) / den
    w2 = ((px - x0) * (y1 - y0) - (py - y0) * (x1 - x0)) / den
    pz = w0 * z0 + w1 * z1 + w2 * z2
    if pz < 0:
        return None

    def seg(ax, ay, bx, by):
        l2 = (bx - ax) ** 2 + (by - ay) ** 2
        if l2 <= O.K_EPS:
            return (px - bx) ** 2 + (py - by) ** 2
        t = min(max(((bx - ax) * (px - ax) + (by - ay) * (py - ay)) / l2, 0.0), 1.0)
        return (ax + t * (bx - ax) - px) ** 2 + (ay + t * (by - ay) - py) ** 2

    d2 = min(seg(x0, y0, x1, y1), seg(x0, y0, x2, y2), seg(x1, y1, x2, y2))
    inside = w0 > 0 and w1 > 0 and w2 > 0
    if not inside and d2 >= O.BLUR_RADIUS:
        return None
    sd = -d2 if inside else d2
    return pz, 1.0 - 1.0 / (1.0 + math.exp(sd / O.SIGMA))


def tile_rasterise(verts, faces, S, K, split="even", list_cap=10 ** 9, bands=1, seed=0):
    """numpy restatement of bin_faces + raster_tile_forward (P0 .. P3), one frame."""
    rects = _rects(verts, faces, S)
    alpha = np.zeros((S, S))
    stats = {"listed": 0, "capped": 0, "passes": 1, "tie_at_cut": 0}
    rng = np.random.default_rng(seed)
    for ty in range((S + TILE - 1) // TILE):
        for tx in range((S + TILE - 1) // TILE):
            x0, y0 = tx * TILE, ty * TILE
            # tile list in ascending face order, tile-local rectangles (bin_fill)
            entries = []
            for f, rc in enumerate(rects):
                if rc is None:
                    continue
                c0, c1, r0, r1 = rc
                if c1 < x0 or c0 > x0 + TILE - 1 or r1 < y0 or r0 > y0 + TILE - 1:
                    continue
                entries.append((f, max(c0 - x0, 0), min(c1 - x0, TILE - 1), max(r0 - y0, 0), min(r1 - y0, TILE - 1)))
            if not entries:
                continue
            for band in range(bands):
                b0, b1 = band * (TILE // bands), (band + 1) * (TILE // bands)
                # contiguous ranges per warp: even, or at random cut points (any contiguous split must do)
                n = len(entries)
                if split == "even":
                    cuts = [min((n + WARPS - 1) // WARPS * w, n) for w in range(WARPS + 1)]
                else:
                    cuts = [0] + sorted(rng.integers(0, n + 1, size=WARPS - 1).tolist()) + [n]
                # P0: corner grids, integrated
                cand = np.zeros((WARPS, TILE, TILE), np.int64)
                for w in range(WARPS):
                    grid = np.zeros((TILE + 1, TILE + 1), np.int64)
                    for (f, c0, c1, r0, r1) in entries[cuts[w]:cuts[w + 1]]:
                        r0c, r1c = max(r0, b0), min(r1, b1 - 1)
                        if r0c > r1c:
                            continue
                        grid[r0c, c0] += 1; grid[r0c, c1 + 1] -= 1; grid[r1c + 1, c0] -= 1; grid[r1c + 1, c1 + 1] += 1
                    cand[w] = grid.cumsum(0).cumsum(1)[:TILE, :TILE]
                    brute = np.zeros((TILE, TILE), np.int64)
                    for (f, c0, c1, r0, r1) in entries[cuts[w]:cuts[w + 1]]:
                        r0c, r1c = max(r0, b0), min(r1, b1 - 1)
                        if r0c <= r1c:
                            brute[r0c:r1c + 1, c0:c1 + 1] += 1
                    assert np.array_equal(cand[w], brute)
                total_c = cand.sum(0)
                listed = total_c > K
                # P0b: list offsets in pixel order (row-major here; any fixed order), passes of list_cap entries
                offs = np.full((TILE, TILE), -1, np.int64)
                run = 0
                for r in range(TILE):
                    for c in range(TILE):
                        if listed[r, c]:
                            offs[r, c] = run
                            run += (total_c[r, c] + 15) // 16 * 16
                n_pass = max(1, -(-run // list_cap))
                stats["passes"] = max(stats["passes"], n_pass)
                stats["listed"] += int(listed[b0:b1].sum())
                P = np.ones((TILE, TILE))
                for p in range(n_pass):
                    active = listed & (offs >= p * list_cap) & (offs < (p + 1) * list_cap)
                    lists = {}
                    cursor = np.zeros((WARPS, TILE, TILE), np.int64)
                    for w in range(WARPS):
                        cursor[w] = offs + cand[:w].sum(0)
                    planes = np.ones((WARPS, TILE, TILE))
                    # P1: every warp sweeps its faces in order
                    for w in range(WARPS):
                        for (f, c0, c1, r0, r1) in entries[cuts[w]:cuts[w + 1]]:
                            for ly in range(max(r0, b0), min(r1, b1 - 1) + 1):
                                for lx in range(c0, c1 + 1):
                                    if listed[ly, lx] and not active[ly, lx]:
                                        continue                         # RT_SKIP
                                    if not listed[ly, lx] and p > 0:
                                        continue
                                    fr = _fragment(verts, faces[f], 1 - (2 * (x0 + lx) + 1) / S, 1 - (2 * (y0 + ly) + 1) / S)
                                    if listed[ly, lx]:
                                        slot = cursor[w, ly, lx]
                                        cursor[w, ly, lx] += 1
                                        assert slot not in lists.setdefault((ly, lx), {})
                                        lists[(ly, lx)][slot] = (fr, f)
                                    elif fr is not None:
                                        planes[w, ly, lx] *= fr[1]
                    # P2a / P2
                    if p == 0:
                        direct = ~listed
                        P[direct] = planes.prod(0)[direct]
                    for (ly, lx), sl in lists.items():
                        slots = sorted(sl)
                        assert slots == list(range(offs[ly, lx], offs[ly, lx] + total_c[ly, lx]))     # dense, own range
                        fids = [sl[s][1] for s in slots]
                        assert fids == sorted(fids)                                                  # face order
                        valid = [(sl[s][0][0], s, sl[s][0][1]) for s in slots if sl[s][0] is not None]
                        valid.sort(key=lambda t: (t[0], t[1]))                                       # (depth, slot)
                        by_fid = sorted(((sl[s][0][0], sl[s][1]) for s in slots if sl[s][0] is not None))
                        assert [v[1] for v in valid[:K]] == [offs[ly, lx] + fids.index(f) for _, f in by_fid[:K]]
                        if len(valid) > K:
                            stats["capped"] += 1
                            stats["tie_at_cut"] += int(valid[K - 1][0] == valid[K][0])
                        prod = 1.0
                        for _, _, m in valid[:K]:
                            prod *= m
                        P[ly, lx] = prod
                for ly in range(b0, b1):
                    for lx in range(TILE):
                        if x0 + lx < S and y0 + ly < S:
                            alpha[y0 + ly, x0 + lx] = 1.0 - P[ly, lx]
    return alpha, stats


@pytest.mark.parametrize("mode", ["even", "random-split", "multi-pass", "bands"])
def test_tile_algorithm_matches_oracle(mode):
    S, K = 48, 6
    verts, faces = _scene(3, 120, S)
    ref = O.soft_silhouette(torch.from_numpy(verts), torch.from_numpy(faces), S, k_faces=K).numpy()
    kw = {"even": {}, "random-split": {"split": "random", "seed": 5}, "multi-pass": {"list_cap": 600},
          "bands": {"bands": 4, "split": "random", "seed": 9}}[mode]
    got, stats = tile_rasterise(verts, faces, S, K, **kw)
    assert int((ref > 0).sum()) > 200 and stats["listed"] > 200 and stats["capped"] > 100 and stats["tie_at_cut"] > 5, stats
    if mode == "multi-pass":
        assert stats["passes"] >= 3, stats
    assert np.abs(got - ref).max() < 1e-12
