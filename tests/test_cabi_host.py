"""CPU: libsmalfit.so loads, exports every symbol include/smalfit.h declares, and the host-side
pieces (model tables, data generator, stage schedule) are consistent.  No compute calls."""
import os
import re

import numpy as np
import pytest
import torch

from smalify_b200 import _cabi, constants as K, model_io, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from smalify_b200 import build
    build.build()
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "smalfit.h")).read()
    declared = set(re.findall(r"SMALFIT_API\s+[\w\s\*]+?\b(smalfit_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smalfit_abi_version() == _cabi.ABI_VERSION


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = model_io.load_asset()
    with pytest.raises(_cabi.SmalfitError):
        _cabi.Handle(c, 0, 2, 64)


def test_fitter_refuses_cpu_device(constants):
    from smalify_b200.smal_fitter import SMALFitter
    blank = (None, torch.zeros(1, 1, 32, 32), torch.zeros(1, 25, 2), torch.zeros(1, 25))
    with pytest.raises(_cabi.SmalfitError):
        SMALFitter("cpu", blank, 1, 1, True, constants=constants)


def test_tables_reconstruct_dense_matrices(constants):
    t = constants.tables
    V = constants.v_template.shape[0]
    W = np.zeros((V, 35), np.float32)
    for v in range(V):
        for k in range(8):
            if t["skin_weight"][v, k] != 0:
                W[v, t["skin_joint"][v, k]] = t["skin_weight"][v, k]
    assert np.array_equal(W, constants.weights)
    W2 = np.zeros_like(W)
    for j in range(35):
        sl = slice(t["skinT_ptr"][j], t["skinT_ptr"][j + 1])
        W2[t["skinT_vert"][sl], j] = t["skinT_weight"][sl]
    assert np.array_equal(W2, constants.weights)
    Jr = np.zeros((V, 35), np.float32)
    for j in range(35):
        sl = slice(t["jreg_ptr"][j], t["jreg_ptr"][j + 1])
        Jr[t["jreg_vert"][sl], j] = t["jreg_weight"][sl]
    assert np.array_equal(Jr, constants.j_regressor)
    # vertex -> face incidence covers every face corner exactly once
    fc = t["v2f_fc"]
    assert len(fc) == constants.faces.size and len(set(fc.tolist())) == len(fc)
    for v in (0, 100, V - 1):
        for e in range(t["v2f_ptr"][v], t["v2f_ptr"][v + 1]):
            assert constants.faces[fc[e] >> 2, fc[e] & 3] == v
    assert np.abs(constants.weights.sum(1) - 1).max() < 1e-5
    assert t["mj_ptr"][-1] == 472 and t["keypoint_joint"].tolist() == list(K.CANONICAL_MODEL_JOINTS)


def test_stage_schedule_and_visibility():
    from smalify_b200.optimize_to_joints import stage_visibility
    assert [r[7] for r in K.STAGE_SCHEDULE] == [150, 400, 600, 800]
    assert [r[8] for r in K.STAGE_SCHEDULE] == [5e-3, 5e-3, 5e-4, 1e-4]
    vis = torch.ones(4, 25)
    v0 = stage_visibility(vis, 0)
    assert v0.sum() == 4 * len(K.TORSO_JOINTS) and torch.equal(stage_visibility(vis, 2), vis)


def test_synthetic_params_are_seeded(constants):
    a = synthetic.ground_truth_params(constants, 5, seed=0)
    b = synthetic.ground_truth_params(constants, 5, seed=0)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["joint_rotations"].abs().max() <= 0.6 + 1e-6
    assert constants.badja_visibility.shape == (201, 25)
