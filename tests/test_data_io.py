"""CPU: ingest / export helpers around the hot path."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from smalify_b200 import constants as K
from smalify_b200 import data_io

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def entries():
    with open(os.path.join(HERE, "golden", "stanford_extra_entries.json")) as f:
        return json.load(f)          # two entries copied from data/StanfordExtra/StanfordExtra_sample.json


def test_rle_decode_matches_annotated_bbox(entries):
    for e in entries:
        h, w = e["img_height"], e["img_width"]
        m = data_io.decode_coco_rle(e["seg"], h, w)
        assert m.shape == (h, w) and set(np.unique(m)) <= {0, 1}
        ys, xs = np.where(m > 0)
        x, y, bw, bh = e["img_bbox"]           # the dataset's own (Stanford Dogs) box of the animal: loose
        tol = 0.03 * max(h, w)
        assert abs(xs.min() - x) <= tol and abs(ys.min() - y) <= tol
        assert abs(xs.max() - (x + bw)) <= tol and abs(ys.max() - (y + bh)) <= tol
        assert 0.25 < m.mean() / (bw * bh / (h * w)) < 1.0     # the animal fills a good part of its box


def test_rle_roundtrip_on_synthetic_mask():
    rng = np.random.default_rng(0)
    m = (rng.random((37, 53)) > 0.6).astype(np.uint8)
    # encode with the inverse of the documented scheme
    flat = m.T.reshape(-1)
    runs, val, cnt = [], 0, 0
    for b in flat:
        if b == val:
            cnt += 1
        else:
            runs.append(cnt); cnt = 1; val ^= 1
    runs.append(cnt)
    out = []
    for i, r in enumerate(runs):
        x = r - runs[i - 2] if i > 2 else r
        more = True
        while more:
            c = x & 0x1F
            x >>= 5
            more = not ((x == 0 and not (c & 0x10)) or (x == -1 and (c & 0x10)))
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    assert np.array_equal(data_io.decode_coco_rle("".join(out), 37, 53), m)
    with pytest.raises(ValueError):
        data_io.decode_coco_rle("".join(out), 36, 53)


def test_stanford_entry_to_fitter_inputs(entries):
    e = [x for x in entries if x["img_path"].endswith("n02099601_176.jpg")][0]      # BASELINE config 1 image
    (rgb, sil, joints, vis), names = data_io.load_stanford_entry(e, 256)
    assert rgb.shape == (1, 3, 256, 256) and sil.shape == (1, 1, 256, 256)
    assert joints.shape == (1, 25, 2) and vis.shape == (1, 25) and names == ["n02099601_176.jpg"]
    assert int(vis.sum()) == 18 and vis[0, 24] == 0                               # 18 visible joints; tail-mid dummy
    s = sil[0, 0].numpy()
    ys, xs = np.where(s > 0)
    # the crop is centred on the silhouette with a 5 % margin (utils.py:19)
    assert 255 - max(xs.max() - xs.min(), ys.max() - ys.min()) < 0.08 * 256
    v = vis[0].bool()
    jj = joints[0][v]
    assert (jj >= 0).all() and (jj <= 256).all()
    # visible joints lie on or near the silhouette
    on = [s[min(255, int(r)), min(255, int(c))] > 0 for r, c in jj.tolist()]
    assert np.mean(on) > 0.6


def test_ply_and_exporter_layout(tmp_path):
    v = np.random.default_rng(1).normal(size=(5, 3)).astype(np.float32)
    f = np.array([[0, 1, 2], [2, 3, 4]])
    p = tmp_path / "m.ply"
    data_io.write_ply(str(p), v, f)
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 5" in head and b"element face 2" in head
    assert len(body) == 5 * 12 + 2 * 13
    assert np.allclose(np.frombuffer(body[:60], dtype="<f4").reshape(5, 3), v)

    class Fake:            # the two methods ResultExporter needs
        constants = type("C", (), {"faces": f})()
        def vertices(self):
            return torch.from_numpy(np.stack([v, v + 1]))
        def export_parameters(self, i):
            return {"global_rotation": np.zeros(3), "joint_rotations": np.zeros((34, 3)), "betas": np.zeros(20),
                    "log_betascale": np.zeros(6), "trans": np.full(3, float(i))}
    ex = data_io.ResultExporter(str(tmp_path / "out"), ["0000.png", "0001.png"])
    ex.stage_id, ex.epoch_name = 10, "0"
    ex.export_fitter(Fake())
    for i in (0, 1):
        d = tmp_path / "out" / f"{i:04}"
        with open(d / "st10_ep0.pkl", "rb") as fh:
            q = pickle.load(fh)
        assert set(q) == {"global_rotation", "joint_rotations", "betas", "log_betascale", "trans"} and q["trans"][0] == i
        assert (d / "st10_ep0.ply").exists()


def _rle_encode(m: np.ndarray) -> str:
    """Inverse of the COCO compressed RLE scheme (column-major runs, delta against the run two back, 5-bit groups)."""
    flat = m.T.reshape(-1)
    runs, val, cnt = [], 0, 0
    for b in flat:
        if b == val:
            cnt += 1
        else:
            runs.append(cnt); cnt = 1; val ^= 1
    runs.append(cnt)
    out = []
    for i, r in enumerate(runs):
        x = r - runs[i - 2] if i > 2 else r
        more = True
        while more:
            c = x & 0x1F
            x >>= 5
            more = not ((x == 0 and not (c & 0x10)) or (x == -1 and (c & 0x10)))
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return "".join(out)


def test_rle_decode_inverts_the_encoder_on_random_masks():
    """Property test (hypothesis): any binary mask of any small shape, including empty, full and single-run masks."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    @settings(max_examples=150, deadline=None)
    @given(hnp.arrays(np.uint8, hnp.array_shapes(min_dims=2, max_dims=2, min_side=1, max_side=40), elements=st.integers(0, 1)),
           st.integers(1, 9))
    def check(mask, blob):
        m = np.repeat(np.repeat(mask, blob, axis=0), blob, axis=1)[:64, :64]      # long runs as well as single pixels
        assert np.array_equal(data_io.decode_coco_rle(_rle_encode(m), m.shape[0], m.shape[1]), m)
    check()
    for m in (np.zeros((7, 5), np.uint8), np.ones((7, 5), np.uint8)):
        assert np.array_equal(data_io.decode_coco_rle(_rle_encode(m), 7, 5), m)
