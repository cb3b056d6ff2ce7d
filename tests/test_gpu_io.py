"""GPU: BASELINE config 1 (the real StanfordExtra image n02099601_176.jpg, WINDOW_SIZE = 1, keypoint
loss only) through ingest -> SMALFitter -> stage-0 fit, and the checkpoint wire format round trip."""
import json
import os

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
