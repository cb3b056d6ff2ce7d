"""GPU: BASELINE config 1 (the real StanfordExtra image n02099601_176.jpg, WINDOW_SIZE = 1, keypoint
loss only) through ingest -> SMALFitter -> stage-0 fit, and the checkpoint wire format round trip."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import smal_oracle as O
from smalify_b200 import constants as K
from smalify_b200 import data_io
from smalify_b200.optimize_to_joints import fit_sequence, stage_visibility

import helpers as H

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def stanford(constants):
    with open(os.path.join(HERE, "golden", "stanford_extra_entries.json")) as f:
        e = [x for x in json.load(f) if x["img_path"].endswith("n02099601_176.jpg")][0]
    return data_io.load_stanford_entry(e, K.CROP_SIZE)


def test_config1_keypoint_only_fit(constants, oracle64, stanford):
    from smalify_b200.smal_fitter import SMALFitter
    data, names = stanford
    rgb, sil, joints, vis = data
    f = SMALFitter("cuda", data, 1, 1, True, constants=constants)
    w = K.STAGE_SCHEDULE[0][:6]
    # first evaluation against the oracle (stage 0: torso joints only, optimize_to_joints.py:98-104)
    v0 = stage_visibility(vis, 0)
    p = O.FitParams.initial(oracle64, 1, K.GLOBAL_ROT_INIT)
    lo, _, go = H.oracle_loss_and_grads(oracle64, p, (rgb, sil, joints, v0), range(1), w, K.CROP_SIZE)
    f.target_visibility = v0.long()
    loss, objs = f([0], w, 0)
    loss.backward()
    assert abs(float(loss) - lo) <= 2e-5 * abs(lo)
    for k in ("global_rotation", "trans"):
        assert H.rel_err(getattr(f, k).grad, go[k]) < 1e-4
    f.target_visibility = vis.long()
    for t in f.parameters():
        t.grad = None
    finals = fit_sequence(f, K.STAGE_SCHEDULE, 1, iters_override=(60, 0, 0, 0))
    assert finals[0] < 0.6 * lo           # 60 of the 150 stage-0 steps already pull the torso keypoints in


def test_checkpoint_wire_format_round_trip(constants, stanford, tmp_path):
    from smalify_b200.smal_fitter import SMALFitter
    data, names = stanford
    f = SMALFitter("cuda", data, 1, 1, True, constants=constants)
    fit_sequence(f, K.STAGE_SCHEDULE, 1, fused=True, iters_override=(5, 5, 0, 0))
    names4 = ["{0:04}.png".format(i) for i in range(1)]         # load_checkpoint reads <dir>/<%04d>/<epoch>.pkl
    ex = data_io.ResultExporter(str(tmp_path), names4)
    ex.stage_id, ex.epoch_name = 10, "0"
    ex.export_fitter(f)
    assert os.path.exists(os.path.join(str(tmp_path), "0000", "st10_ep0.ply"))
    g = SMALFitter("cuda", data, 1, 1, True, constants=constants)
    g.load_checkpoint(str(tmp_path), "st10_ep0")
    for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans"):
        assert torch.allclose(getattr(g, k).detach(), getattr(f, k).detach(), atol=1e-7), k
    a0, _ = f.render()
    a1, _ = g.render()
    assert torch.equal(a0, a1)


def test_script_main_and_generate_video(constants, stanford, tmp_path):
    """optimize_to_joints.main / generate_video.main equivalents on the config-1 image: a short schedule, the
    collage + pkl + ply exports under the reference's names, then the reload -> per-frame collage path."""
    import cv2
    from smalify_b200.optimize_to_joints import RunConfig, generate_video, main
    data, _ = stanford
    names = ["0000.jpg"]                         # load_checkpoint addresses frames as <dir>/<%04d>/<epoch>.pkl
    cfg = RunConfig(OUTPUT_DIR=str(tmp_path / "checkpoints" / "run"), SEQUENCE_OR_IMAGE_NAME="stanfordextra:x", WINDOW_SIZE=1,
                    VIS_FREQUENCY=4, CHECKPOINT_NAME="run")
    model, finals = main(cfg, constants=constants, data=(data, names), iters_override=(6, 5, 0, 0))
    d = os.path.join(cfg.OUTPUT_DIR, "0000")
    for stem in ("st0_ep0", "st0_ep4", "st1_ep4", "st10_ep0"):
        assert os.path.exists(os.path.join(d, stem + ".png")) and os.path.exists(os.path.join(d, stem + ".pkl")), stem
    S = K.CROP_SIZE
    assert cv2.imread(os.path.join(d, "st10_ep0.png")).shape == (S, 5 * S, 3)
    assert finals[1] is not None and finals[1] == finals[1]
    out = generate_video(cfg, constants=constants, data=(data, names), checkpoints_root=str(tmp_path / "checkpoints"),
                         export_root=str(tmp_path / "exported"))
    img = cv2.imread(os.path.join(out, "0000.png"))
    assert img is not None and img.shape == (S, 5 * S, 3)
    # the reloaded parameters are the exported ones: same collage up to the png quantisation
    ref = cv2.imread(os.path.join(d, "st10_ep0.png"))
    assert float(np.mean(np.abs(img.astype(np.int32) - ref.astype(np.int32)) > 2)) < 0.01


def test_checkpoint_round_trip_with_one_shape_per_frame(constants, oracle64, tmp_path):
    """Every frame has its own shape (BASELINE config 4): the per-frame pickle holds that frame's (20,) betas and
    (6,) log scales (the reference's layout, smal_fitter.py:213-219), load_checkpoint restores them per frame and
    writes into the existing parameter storage -- a FusedFit created before keeps working on the loaded values."""
    import pickle
    from smalify_b200 import synthetic
    from smalify_b200.smal_fitter import FusedFit, SMALFitter
    S, n = 64, 3
    data, gt = synthetic.make_subsequence(constants, n, range(n), S, H.oracle_renderer(oracle64, S), seed=2)
    f = SMALFitter("cuda", data, 1, 1, True, constants=constants, per_frame_shapes=True)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        f.betas += 0.1 * torch.randn(n, 20, generator=g).to(f.device)
        f.log_beta_scales += 0.05 * torch.randn(n, 6, generator=g).to(f.device)
        f.trans += 0.02 * torch.randn(n, 3, generator=g).to(f.device)
    for i in range(n):
        d = f.export_parameters(i)
        assert d["betas"].shape == (20,) and d["log_betascale"].shape == (6,) and d["joint_rotations"].shape == (34, 3)
        os.makedirs(tmp_path / "{0:04}".format(i), exist_ok=True)
        with open(tmp_path / "{0:04}".format(i) / "st10_ep0.pkl", "wb") as fh:
            pickle.dump(d, fh)
    h = SMALFitter("cuda", data, 1, 1, True, constants=constants, per_frame_shapes=True)
    loop = FusedFit(h, 1)                                  # re-points the parameters at its flat buffer
    ptrs = [p.data_ptr() for p in h.parameters()]
    h.load_checkpoint(str(tmp_path), "st10_ep0")
    assert ptrs == [p.data_ptr() for p in h.parameters()]
    for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans"):
        assert torch.equal(getattr(h, k).detach(), getattr(f, k).detach()), k
    # the kernels see the loaded values: same loss from both fitters, and the fused loop steps without a fault
    w = K.STAGE_SCHEDULE[1][:6]
    la, _ = f(list(range(n)), w, 1)
    lb, _ = h(list(range(n)), w, 1)
    assert float(la) == float(lb)
    loop.step(w, 0.0, 1e-3)
    torch.cuda.synchronize()
    h.check_faults()
