"""CPU: host-side pieces around the path that need no GPU -- joint markers, the exporter's file layout,
run settings, the limits table."""
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from smalify_b200 import constants as K
from smalify_b200 import data_io, visualization
from smalify_b200.optimize_to_joints import FrameExporter, RunConfig, stage_visibility


def test_draw_joints_markers_and_parking():
    """draw_smal_joints.py:9-46: one marker per (row, col) landmark in the joint's colour; invisible joints are parked
    along the top edge, 10 px apart."""
    img = torch.ones(1, 3, 64, 64)
    lm = torch.zeros(1, K.N_KEYPOINTS, 2)
    lm[0, :, 0] = 40.0                              # row
    lm[0, :, 1] = torch.arange(K.N_KEYPOINTS) * 2.0 + 5.0
    vis = torch.ones(1, K.N_KEYPOINTS)
    vis[0, 3] = 0
    vis[0, 7] = 0
    out = visualization.draw_joints(img, lm, vis)
    assert out.shape == (1, 3, 64, 64) and 0.0 <= float(out.min()) and float(out.max()) <= 1.0
    arr = (out[0].permute(1, 2, 0).numpy() * 255).round().astype(int)
    assert (arr[40, 5] == np.array(visualization.MARKER_COLORS[0])).all()          # centre of joint 0's marker
    # joints 3 and 7 are parked at (x, y) = (0, 0) and (10, 0) in their own colours
    assert (arr[0, 0] == np.array(visualization.MARKER_COLORS[3])).all()
    assert (arr[0, 10] == np.array(visualization.MARKER_COLORS[7])).all()
    assert len(visualization.MARKER_COLORS) == len(visualization.MARKER_TYPE) == K.N_KEYPOINTS


def test_exporters_write_reference_layout(tmp_path):
    """optimize_to_joints.py:25-53 and generate_video.py:26-36."""
    import cv2
    ex = data_io.ResultExporter(str(tmp_path / "ckpt"), ["a.jpg", "dir_b.png"])
    assert [os.path.basename(d) for d in ex.output_dirs] == ["a", "dir_b"]
    ex.stage_id, ex.epoch_name = 2, "300"
    collage = np.zeros((8, 40, 3), np.uint8)
    collage[:, :, 0] = 255                          # RGB red
    verts = torch.rand(2, 5, 3)
    faces = np.array([[0, 1, 2], [2, 3, 4]])
    params = {"betas": np.zeros(20), "trans": np.ones(3)}
    ex.export(collage, 1, 1, params, verts, faces)
    stem = os.path.join(ex.output_dirs[1], "st2_ep300")
    back = cv2.imread(stem + ".png")
    assert back.shape == (8, 40, 3) and (back[0, 0] == [0, 0, 255]).all()           # cv2 reads BGR
    with open(stem + ".pkl", "rb") as f:
        assert set(pkl.load(f)) == {"betas", "trans"}
    raw = open(stem + ".ply", "rb").read()
    head = raw[:raw.index(b"end_header\n")].decode("ascii").splitlines()
    assert head[0] == "ply" and "element vertex 5" in head and "element face 2" in head
    assert len(raw) == raw.index(b"end_header\n") + len(b"end_header\n") + 5 * 12 + 2 * 13        # binary body
    fe = FrameExporter(str(tmp_path / "exported"))
    fe.export(collage, 0, 12, params, verts, faces)
    assert os.path.exists(os.path.join(fe.output_dir, "0012.png")) and os.path.exists(os.path.join(fe.output_dir, "0012.pkl"))


def test_run_config_and_stage_visibility():
    cfg = RunConfig(WINDOW_SIZE=4, SEQUENCE_OR_IMAGE_NAME="stanfordextra:x.jpg")
    assert cfg.WINDOW_SIZE == 4 and cfg.CROP_SIZE == K.CROP_SIZE and cfg.EPOCH_NAME == "st10_ep0"
    assert len(cfg.OPT_SCHEDULE) == 4 and cfg.OPT_SCHEDULE[0][7] == 150
    with pytest.raises(TypeError):
        RunConfig(WINDOWSIZE=4)
    v = torch.ones(2, K.N_KEYPOINTS)
    v0 = stage_visibility(v, 0)
    assert set(torch.nonzero(v0[0]).flatten().tolist()) == set(K.TORSO_JOINTS)        # optimize_to_joints.py:98-104
    assert torch.equal(stage_visibility(v, 1), v)


def test_joint_limits_table_shape():
    lo, hi = K.joint_limits()
    assert lo.shape == hi.shape == (K.N_POSE, 3) and np.all(lo <= hi)
    assert np.isfinite(lo[:32]).all() and np.isfinite(hi[:32]).all() and not np.isfinite(lo[32:]).any()


def test_subsequence_is_a_slice_of_the_sequence(constants, oracle64):
    """synthetic.make_subsequence: a rank that builds only its shard of the seeded sequence gets exactly the frames
    (targets, keypoint noise, visibility rows, ground truth) the whole sequence has at those positions -- also in the
    padded layout a frame-sharded fitter takes."""
    import torch
    import helpers as H
    from smalify_b200 import synthetic
    S, n, idx = 24, 5, [1, 3, 4]
    render = H.oracle_renderer(oracle64, S)
    (rgb, sil, joints, vis), gt = synthetic.make_sequence(constants, n, S, render, seed=0)
    (rgb2, sil2, joints2, vis2), gt2 = synthetic.make_subsequence(constants, n, idx, S, render, seed=0)
    assert torch.equal(sil2, sil[idx]) and torch.equal(joints2, joints[idx]) and torch.equal(vis2, vis[idx])
    assert torch.equal(gt2["global_rotation"], gt["global_rotation"][idx]) and torch.equal(gt2["betas"], gt["betas"])
    (_, sil3, joints3, vis3), _ = synthetic.make_subsequence(constants, n, [3, 4], S, render, seed=0, pad_to=(3, n))
    assert sil3.shape[0] == n and torch.equal(sil3[3:5], sil[3:5]) and float(sil3[:3].abs().sum()) == 0.0
    assert torch.equal(joints3[3:5], joints[3:5]) and torch.equal(vis3[3:5], vis[3:5])
    # one shape per frame: the ground truth carries a row per frame
    (_, sil4, _, _), gt4 = synthetic.make_subsequence(constants, n, [0, 2], S, render, seed=0, per_frame_shapes=True)
    assert gt4["betas"].shape == (2, 20) and sil4.shape[0] == 2
