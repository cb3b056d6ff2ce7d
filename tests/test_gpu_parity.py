"""GPU parity: libsmalfit (through the SMALFitter / C-ABI surface) against the CPU oracle
on identical seeded inputs.  Tolerances: loss terms rel 1e-5-class (fp32 kernels
vs fp64 oracle: 2e-5), gradients 1e-4 of the tensor's max magnitude, silhouettes 2e-5 abs."""
import numpy as np
import pytest
import torch

from oracle import smal_oracle as O
from smalify_b200 import constants as K
from smalify_b200 import synthetic

import helpers as H

pytestmark = pytest.mark.gpu

S_SMALL = 64
N_SMALL = 3
STAGE1 = K.STAGE_SCHEDULE[1][:6]
STAGE0 = K.STAGE_SCHEDULE[0][:6]


@pytest.fixture(scope="module")
def seq(constants, oracle64):
    data, gt = synthetic.make_sequence(constants, N_SMALL, S_SMALL, H.oracle_renderer(oracle64, S_SMALL), seed=0)
    return data, gt


@pytest.fixture(scope="module")
def fitter(constants, seq):
    from smalify_b200.smal_fitter import SMALFitter
    data, _ = seq
    return SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants)


def _states(oracle64, gt):
    init = O.FitParams.initial(oracle64, N_SMALL, K.GLOBAL_ROT_INIT)
    mid = H.perturbed_params(oracle64, gt, seed=5)
    near = H.perturbed_params(oracle64, gt, seed=6, scale=0.2)
    return {"init": init, "mid": mid, "near": near}


def test_vertices_and_keypoints(fitter, oracle64, seq):
    _, gt = seq
    for name, p in _states(oracle64, gt).items():
        H.load_params_into(fitter, p)
        v = fitter.vertices().cpu().double()
        _, kp = fitter.render()
        theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
        vo, jo, _ = O.smal_forward(oracle64, p.betas.expand(N_SMALL, 20), theta, p.log_beta_scales.expand(N_SMALL, 6))
        vo = vo + p.trans[:, None]
        jo = jo + p.trans[:, None]
        assert (v - vo).abs().max() < 2e-6, name
        kpo = O.project_points_screen(jo[:, list(O.CANONICAL)], S_SMALL)
        assert (kp.cpu().double() - kpo).abs().max() < 2e-4, name      # pixels


def test_silhouette(fitter, oracle64, seq):
    _, gt = seq
    for name, p in _states(oracle64, gt).items():
        H.load_params_into(fitter, p)
        alpha, _ = fitter.render()
        theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
        vo, _, _ = O.smal_forward(oracle64, p.betas.expand(N_SMALL, 20), theta, p.log_beta_scales.expand(N_SMALL, 6))
        ao = O.render_silhouettes(oracle64, vo + p.trans[:, None], S_SMALL)
        err = (alpha.cpu().double() - ao).abs()
        # fp32 pz can reorder two fragments at the K=100 cut in a handful of pixels
        assert float((err > 2e-5).double().mean()) < 2e-3, (name, float(err.max()))
        assert float(err.mean()) < 1e-5, name
    c = fitter.counters()
    assert c["capped_pixels"] > 0 and c["dropped_bin_entries"] == 0


@pytest.mark.parametrize("weights,label", [(STAGE0, "stage0"), (STAGE1, "stage1"), (K.STAGE_SCHEDULE[2][:6], "stage2")])
def test_loss_and_gradients(fitter, oracle64, seq, weights, label):
    data, gt = seq
    for name, p in _states(oracle64, gt).items():
        lo, objs_o, go = H.oracle_loss_and_grads(oracle64, p, data, range(N_SMALL), weights, S_SMALL)
        H.load_params_into(fitter, p)
        for t in fitter.parameters():
            t.grad = None
            t.requires_grad_(True)
        loss, objs = fitter(list(range(N_SMALL)), weights, 1)
        loss.backward()
        assert abs(float(loss) - lo) <= 2e-5 * abs(lo) + 1e-6, (label, name, float(loss), lo)
        for k, v in objs_o.items():
            assert abs(float(objs[k]) - v) <= 3e-5 * abs(v) + 1e-6, (label, name, k, float(objs[k]), v)
        for k in ("global_rotation", "trans", "joint_rotations", "betas", "log_beta_scales"):
            g = getattr(fitter, k).grad
            if float(go[k].abs().max()) == 0.0:
                assert g is None or float(g.abs().max()) < 1e-6
                continue
            assert H.rel_err(g, go[k]) < 1e-4, (label, name, k, H.rel_err(g, go[k]))


@pytest.mark.parametrize("knobs", [
    {"SMALFIT_RT_LISTCAP": "2048"},                                    # tiles finished in several passes (RT_SKIP path)
    {"SMALFIT_RT_NSUB": "8", "SMALFIT_RT_SPLITLEN": "64"},            # every tile cut into 8 row bands
    {"SMALFIT_RT_LISTCAP": "1024", "SMALFIT_RT_NSUB": "2", "SMALFIT_RT_SPLITLEN": "64"},
])
def test_tile_rasteriser_paths_agree(constants, fitter, oracle64, seq, knobs, monkeypatch):
    """The tile rasteriser's rarely taken paths (multi-pass tiles, row-band items), forced through the
    create-time knobs, against the default configuration: same silhouettes,
    loss and gradients (selection is exact in all of them; products are grouped differently: 1e-6)."""
    from smalify_b200.smal_fitter import SMALFitter
    data, gt = seq
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    other = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants)
    for k in knobs:
        monkeypatch.delenv(k)
    for name, p in _states(oracle64, gt).items():
        res = []
        for f in (fitter, other):
            H.load_params_into(f, p)
            for t in f.parameters():
                t.grad = None
                t.requires_grad_(True)
            loss, _ = f(list(range(N_SMALL)), STAGE1, 1)
            loss.backward()
            alpha, _ = f.render()
            res.append((float(loss), alpha.clone(), {k: getattr(f, k).grad.clone() for k in ("global_rotation", "trans", "joint_rotations", "betas")}))
        (la, aa, ga), (lb, ab, gb) = res
        assert abs(la - lb) <= 2e-6 * abs(la), (name, la, lb)
        assert float((aa - ab).abs().max()) < 2e-6, name
        for k in ga:
            assert H.rel_err(gb[k], ga[k].double().cpu()) < 2e-5, (name, k)


def test_joint_limit_term(constants, oracle64, seq):
    """Row 8f-4: the optional joint-limit hinge (reference: commented out at smal_fitter.py:146-151, limits of
    priors/joint_limits_prior.py) against the oracle, value and gradient, with many joints past their limits;
    off by default: w_limit is then ignored exactly as the reference ignores it."""
    from smalify_b200.smal_fitter import SMALFitter
    data, gt = seq
    f = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants, joint_limits=True)
    limits = K.joint_limits()
    p = H.perturbed_params(oracle64, gt, seed=9)
    gen = torch.Generator().manual_seed(4)
    p.joint_rotations = p.joint_rotations + 0.8 * torch.randn(p.joint_rotations.shape, generator=gen, dtype=p.joint_rotations.dtype)
    for weights in (STAGE1, (0.0, 0.0, 0.0, 0.0, 100.0, 0.0)):
        lo, objs_o, go = H.oracle_loss_and_grads(oracle64, p, data, range(N_SMALL), weights, S_SMALL, joint_limits=limits)
        assert objs_o["limit"] > 0
        H.load_params_into(f, p)
        for t in f.parameters():
            t.grad = None
            t.requires_grad_(True)
        loss, objs = f(list(range(N_SMALL)), weights, 1)
        loss.backward()
        assert abs(float(objs["limit"]) - objs_o["limit"]) <= 2e-6 * objs_o["limit"]
        assert abs(float(loss) - lo) <= 2e-5 * abs(lo)
        assert H.rel_err(f.joint_rotations.grad, go["joint_rotations"]) < 1e-4
    # default fitter: same weights, no limit term
    f.set_joint_limits(None)
    loss2, objs2 = f(list(range(N_SMALL)), STAGE1, 1)
    lo2, objs_o2, _ = H.oracle_loss_and_grads(oracle64, p, data, range(N_SMALL), STAGE1, S_SMALL)
    assert "limit" not in objs2 and abs(float(loss2) - lo2) <= 2e-5 * abs(lo2)


def test_focal_parameter(constants, oracle64, seq):
    """Row 8f-4: a focal parameter (the reference camera is fixed at fov 60 degrees, p3d_renderer.py:22-23).
    Silhouettes, keypoints, loss and every gradient including dL/dfocal against the oracle at a non-default
    focal; dL/dfocal also against a central finite difference of the GPU loss."""
    from smalify_b200.smal_fitter import SMALFitter
    data, gt = seq
    f0 = 1.55
    f = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants, focal=f0)
    assert len(list(f.parameters())) == 6 and not f.focal.requires_grad
    f.focal.requires_grad_(True)
    p = H.perturbed_params(oracle64, gt, seed=5)
    for weights in (STAGE1, STAGE0):
        foc = torch.tensor(f0, dtype=torch.float64)
        lo, objs_o, go = H.oracle_loss_and_grads(oracle64, p, data, range(N_SMALL), weights, S_SMALL, focal=foc)
        H.load_params_into(f, p)
        for t in f.parameters():
            t.grad = None
        loss, objs = f(list(range(N_SMALL)), weights, 1)
        loss.backward()
        assert abs(float(loss) - lo) <= 2e-5 * abs(lo), (float(loss), lo)
        for k in ("global_rotation", "trans", "joint_rotations"):
            assert H.rel_err(getattr(f, k).grad, go[k]) < 1e-4, k
        gf = float(f.focal.grad)
        assert abs(gf - float(go["focal"])) <= 1e-4 * abs(float(go["focal"])) + 1e-5, (gf, float(go["focal"]))
    # the silhouette the backward differentiates really is rendered with this focal
    alpha, kp = f.render()
    theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
    vo, jo, _ = O.smal_forward(oracle64, p.betas.expand(N_SMALL, 20), theta, p.log_beta_scales.expand(N_SMALL, 6))
    ao = O.render_silhouettes(oracle64, vo + p.trans[:, None], S_SMALL, focal=f0)
    assert float((alpha.cpu().double() - ao).abs().mean()) < 1e-5
    kpo = O.project_points_screen((jo + p.trans[:, None])[:, list(O.CANONICAL)], S_SMALL, focal=f0)
    assert (kp.cpu().double() - kpo).abs().max() < 2e-4
    # default construction: five parameters, reference camera
    assert len(list(SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants).parameters())) == 5


def test_windows_and_temporal(fitter, oracle64, seq):
    data, gt = seq
    p = H.perturbed_params(oracle64, gt, seed=9)
    w = STAGE1
    for t in p.tensors():
        t.requires_grad_(True)
        t.grad = None
    rgb, sil, joints, vis = data
    total = O.epoch_loss(oracle64, p, sil, joints, vis, 2, w, 100.0, S_SMALL)      # windows of 2 + 1 frames
    total.backward()
    H.load_params_into(fitter, p)
    for t in fitter.parameters():
        t.grad = None
        t.requires_grad_(True)
    acc = 0
    for j in range(0, N_SMALL, 2):
        loss, _ = fitter(list(range(j, min(N_SMALL, j + 2))), w, 1)
        acc = acc + loss.mean()
    jl, gl, tl = fitter.get_temporal(100.0)
    acc = acc + jl + gl + tl
    acc.backward()
    assert abs(float(acc) - float(total)) <= 2e-5 * abs(float(total))
    for k in ("global_rotation", "trans", "joint_rotations", "betas", "log_beta_scales"):
        assert H.rel_err(getattr(fitter, k).grad, getattr(p, k).grad) < 1e-4, k


def test_per_frame_shapes_equal_independent_fits(constants, oracle64, seq):
    """Extension (BASELINE config 4): one shape per frame = a batch of independent single-image problems.
    Oracle side: the sum of single-frame losses, each frame with its own betas / log-scales."""
    from smalify_b200.smal_fitter import SMALFitter
    data, gt = seq
    rgb, sil, joints, vis = data
    f = SMALFitter("cuda", data, 1, 1, True, constants=constants, per_frame_shapes=True)
    g = torch.Generator().manual_seed(21)
    frames = []
    total = torch.zeros((), dtype=torch.float64)
    w = STAGE1
    for i in range(N_SMALL):
        p = H.perturbed_params(oracle64, {k: (v[i:i + 1] if v.dim() > 1 else v) for k, v in gt.items()}, seed=30 + i)
        for t in p.tensors():
            t.requires_grad_(True)
        one = (None, sil[i:i + 1], joints[i:i + 1], vis[i:i + 1])
        loss, _ = O.fitter_forward(oracle64, p, one[1], one[2], one[3], range(1), w, S_SMALL)
        total = total + loss
        frames.append(p)
    total.backward()
    with torch.no_grad():
        for i, p in enumerate(frames):
            f.betas[i] = p.betas.float().to(f.device)
            f.log_beta_scales[i] = p.log_beta_scales.float().to(f.device)
            f.global_rotation[i] = p.global_rotation[0].float().to(f.device)
            f.joint_rotations[i] = p.joint_rotations[0].float().to(f.device)
            f.trans[i] = p.trans[0].float().to(f.device)
    loss, _ = f(list(range(N_SMALL)), w, 1)
    loss.backward()
    assert abs(float(loss) - float(total)) <= 2e-5 * abs(float(total))
    for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans"):
        ref = torch.stack([getattr(p, k).grad.reshape(getattr(f, k).shape[1:]) for p in frames])
        assert H.rel_err(getattr(f, k).grad, ref) < 1e-4, k


@pytest.mark.parametrize("S", [40, 96])
def test_edge_cases_clipping_behind_camera_empty_targets(constants, oracle64, S):
    """Image side not a multiple of the 32-pixel tile; one animal half outside the image, one so close
    that part of the mesh is behind the camera plane (z < 0 faces / fragments are dropped), an empty
    target mask and a frame without visible keypoints."""
    from smalify_b200.smal_fitter import SMALFitter
    N = 3
    m = oracle64
    gt = synthetic.ground_truth_params(constants, N, seed=4)
    gt["trans"] = torch.tensor([[0.95, 0.3, 0.4], [0.0, 0.1, 2.45], [-0.2, -1.1, 0.9]])
    data, _ = synthetic.make_sequence(constants, N, S, H.oracle_renderer(m, S), seed=4)
    rgb, sil, joints, vis = data
    sil = sil.clone()
    sil[2] = 0.0                      # empty target
    vis = vis.clone()
    vis[1] = 0.0                      # nothing visible
    data = (rgb, sil, joints, vis)
    p = O.FitParams(global_rotation=gt["global_rotation"].double(), joint_rotations=gt["joint_rotations"].double(),
                    betas=gt["betas"].double(), log_beta_scales=gt["log_beta_scales"].double(), trans=gt["trans"].double())
    w = K.STAGE_SCHEDULE[2][:6]
    lo, objs_o, go = H.oracle_loss_and_grads(m, p, data, range(N), w, S)
    f = SMALFitter("cuda", data, N, 1, True, constants=constants)
    H.load_params_into(f, p)
    loss, objs = f(list(range(N)), w, 2)
    loss.backward()
    assert torch.isfinite(loss)
    assert abs(float(loss) - lo) <= 3e-5 * abs(lo), (float(loss), lo)
    for k in ("global_rotation", "trans", "joint_rotations", "betas", "log_beta_scales"):
        g = getattr(f, k).grad
        assert torch.isfinite(g).all()
        assert H.rel_err(g, go[k]) < 1e-4, (k, H.rel_err(g, go[k]))
    alpha, _ = f.render()
    theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
    vo, _, _ = O.smal_forward(m, p.betas.expand(N, 20), theta, p.log_beta_scales.expand(N, 6))
    ao = O.render_silhouettes(m, vo + p.trans[:, None], S)
    err = (alpha.cpu().double() - ao).abs()
    assert float((err > 5e-5).double().mean()) < 5e-3, float(err.max())
    assert f.counters()["dropped_bin_entries"] == 0


def test_c_abi_rejects_bad_arguments(fitter):
    """Error behaviour of the boundary: negative return codes + message, never a crash."""
    import ctypes
    from smalify_b200 import _cabi
    from smalify_b200.smal_fitter import _tensors, _ptr, _stream, _weights6
    h = fitter._handle
    params = _tensors(fitter.betas, fitter.log_beta_scales, fitter.global_rotation, fitter.joint_rotations, fitter.trans)
    terms = torch.zeros(8, device=fitter.device)
    rc = h.lib.smalfit_loss_grad(h.h, ctypes.byref(params), 0, N_SMALL + 1, _weights6(STAGE1), 1, None, _ptr(terms), _stream(fitter.device))
    assert rc == -1 and b"bad arguments" in h.lib.smalfit_last_error(h.h)
    bad = _tensors(None, fitter.log_beta_scales, fitter.global_rotation, fitter.joint_rotations, fitter.trans)
    rc = h.lib.smalfit_loss_grad(h.h, ctypes.byref(bad), 0, 1, _weights6(STAGE1), 1, None, _ptr(terms), _stream(fitter.device))
    assert rc == -1 and b"NULL" in h.lib.smalfit_last_error(h.h)
    with pytest.raises(ValueError):
        fitter([0, 2], STAGE1, 1)          # non-contiguous window
    with pytest.raises(_cabi.SmalfitError):
        _cabi.Handle(fitter.constants, fitter.device.index, 4, 2048)      # image side above the supported 1024


def test_missing_library_fails_loudly(tmp_path):
    from smalify_b200 import _cabi
    with pytest.raises(_cabi.SmalfitError):
        _cabi.load_library(str(tmp_path / "libsmalfit_missing.so"))


def test_fused_step_equals_the_unfused_sequence(constants, seq):
    """smalfit_fused_step (temporal term folded into the frame kernels, Adam in the step-tail kernel) against the
    call sequence it replaces (smalfit_loss_grad, smalfit_temporal, smalfit_adam_step): the same operations in the
    same order, so parameters, Adam state and gradients agree BIT FOR BIT after several steps; the loss terms to
    rounding (their sums are grouped differently)."""
    import ctypes
    from smalify_b200 import _cabi
    from smalify_b200.smal_fitter import FusedFit, SMALFitter, _ptr, _stream, _weights6
    data, _ = seq
    row = K.STAGE_SCHEDULE[1]
    w, w_temp, lr = row[:6], row[6], row[8]
    fa = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants)
    fb = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants)
    la, lb = FusedFit(fa, 2), FusedFit(fb, 2)           # windows of 2 + 1 frames
    assert la.fused_tail
    h = fb._handle
    st = _stream(fb.device)
    temporal = torch.zeros(3, device=fb.device)
    for step in range(1, 6):
        la.step(w, w_temp, lr, use_graph=(step > 2))
        params, grads, m, v = lb._t(0), lb._t(1), lb._t(2), lb._t(3)
        h.check(h.lib.smalfit_loss_grad(h.h, ctypes.byref(params), 0, N_SMALL, _weights6(w), lb.n_windows, ctypes.byref(grads),
                                        _ptr(lb.terms), st), "loss_grad")
        h.check(h.lib.smalfit_temporal(h.h, ctypes.byref(params), N_SMALL, float(w_temp), ctypes.byref(grads), _ptr(temporal), st), "temporal")
        tr = (ctypes.c_int32 * 5)(1, 1, 1, 1, 1)
        h.check(h.lib.smalfit_adam_step(h.h, ctypes.byref(params), ctypes.byref(grads), ctypes.byref(m), ctypes.byref(v), N_SMALL, tr,
                                        float(lr), K.ADAM_BETAS[0], K.ADAM_BETAS[1], K.ADAM_EPS, step, st), "adam")
        torch.cuda.synchronize()
        total = sum(la.sizes)
        assert torch.equal(la.flat_g[:total], lb.flat_g[:total]), step
        assert torch.equal(la.flat_p, lb.flat_p) and torch.equal(la.flat_m, lb.flat_m) and torch.equal(la.flat_v, lb.flat_v), step
        tb = float(lb.terms[_cabi.L_TOTAL] + temporal.sum())
        assert abs(float(la.total_loss()) - tb) <= 1e-6 * abs(tb)
        assert torch.allclose(la.temporal_terms, temporal, rtol=1e-6, atol=0)
    assert fa._handle.status() == 0


def test_pool_overflow_is_reported_not_silent(constants, seq):
    """A (face, tile) pool that is too small must not pass silently (the dropped entries make the silhouette terms
    inexact): the sticky fault shows in smalfit_status and the next hot-path call fails."""
    from smalify_b200 import _cabi
    from smalify_b200.smal_fitter import SMALFitter
    data, _ = seq
    f = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants, pool_entries_per_frame=1024)
    loss, _ = f(list(range(N_SMALL)), STAGE1, 1)
    torch.cuda.synchronize()
    assert f._handle.status() & _cabi.STATUS_POOL_OVERFLOW
    assert f.counters()["dropped_bin_entries"] > 0
    with pytest.raises(_cabi.SmalfitError):
        f(list(range(N_SMALL)), STAGE1, 1)
    with pytest.raises(_cabi.SmalfitError):
        f.check_faults()


def test_frame_shard_handle_matches_full_handle(constants, seq):
    """A handle created for a shard of the frames (workspace and targets sized for them only) gives, on its frames,
    the gradients of the full handle; frames outside the shard are rejected."""
    from smalify_b200 import _cabi
    from smalify_b200.smal_fitter import SMALFitter
    data, gt = seq
    full = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants)
    part = SMALFitter("cuda", data, N_SMALL, 1, True, constants=constants, frame_shard=(1, 3))
    res = []
    for f in (full, part):
        for t in f.parameters():
            t.grad = None
            t.requires_grad_(True)
        f.set_windows(np.full(N_SMALL, 2, np.int32))
        f._windows_for = np.full(N_SMALL, 2, np.int32)
        loss, _ = f([1, 2], STAGE1, 1)
        loss.backward()
        res.append((float(loss), {k: getattr(f, k).grad.clone() for k in ("global_rotation", "trans", "joint_rotations", "betas", "log_beta_scales")}))
    assert res[0][0] == res[1][0]
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k
    with pytest.raises(_cabi.SmalfitError):
        part([0, 1], STAGE1, 1)


def test_staged_targets_equal_serial_uploads(constants, oracle64, seq):
    """Double-buffered targets (smalfit_stage_targets on a copy stream + smalfit_swap_targets) against the serial
    smalfit_set_targets of the same target sequence: a fused loop whose targets alternate between two sets every step
    ends with BIT-IDENTICAL parameters and Adam state either way, eager or graph-replayed (one graph per set), also
    on a frame-shard handle; swapping with nothing staged is an error."""
    from smalify_b200 import _cabi
    from smalify_b200.smal_fitter import FusedFit, SMALFitter, _ptr, _stream
    data_a, _ = seq
    data_b, _ = synthetic.make_sequence(constants, N_SMALL, S_SMALL, H.oracle_renderer(oracle64, S_SMALL), seed=3)
    row = K.STAGE_SCHEDULE[1]
    w, w_temp, lr = row[:6], row[6], row[8]

    def host_set(data):
        _, sil, joints, vis = data
        return ((sil.reshape(N_SMALL, S_SMALL, S_SMALL) > 0.5).to(torch.uint8).contiguous().pin_memory(),
                joints.reshape(N_SMALL, K.N_KEYPOINTS, 2).float().contiguous().pin_memory(),
                vis.reshape(N_SMALL, K.N_KEYPOINTS).to(torch.uint8).contiguous().pin_memory())
    sets = [host_set(data_a), host_set(data_b)]
    assert not torch.equal(sets[0][0], sets[1][0])

    for shard in (None, (1, 3)):
        lo, hi = shard or (0, N_SMALL)
        fa = SMALFitter("cuda", data_a, N_SMALL, 1, True, constants=constants, frame_shard=shard)
        fb = SMALFitter("cuda", data_a, N_SMALL, 1, True, constants=constants, frame_shard=shard)
        with pytest.raises(_cabi.SmalfitError):
            fa.swap_targets()
        la, lb = FusedFit(fa, N_SMALL, frame_shard=shard), FusedFit(fb, N_SMALL, frame_shard=shard)
        copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        h = fb._handle
        seen = set()
        for step in range(8):
            sil, joints, vis = (t[lo:hi] for t in sets[step & 1])
            # pipelined: stage on the copy stream (after the step that last read that set), swap, step
            done = torch.cuda.Event()
            done.record(main)
            copy_stream.wait_event(done)
            ev = fa.stage_targets(sil, joints, vis, copy_stream)
            seen.add(fa.swap_targets(ev))
            la.step(w, w_temp, lr, use_graph=(step >= 2))
            # serial: overwrite the one set in place on the main stream
            h.check(h.lib.smalfit_set_targets(h.h, lo, hi - lo, _ptr(sil), _ptr(joints), _ptr(vis), 1, _stream(fb.device)), "set_targets")
            lb.step(w, w_temp, lr, use_graph=(step >= 2))
            torch.cuda.synchronize()
            assert torch.equal(la.flat_p, lb.flat_p), (shard, step)
            assert torch.equal(la.flat_m, lb.flat_m) and torch.equal(la.flat_v, lb.flat_v), (shard, step)
            assert float(la.total_loss()) == float(lb.total_loss()), (shard, step)
        assert seen == {0, 1}
        assert len(la._graphs) == 2 and len(lb._graphs) == 1
        assert fa._handle.status() == 0
    # the drop-in surface reads the swapped-in set as well
    fc = SMALFitter("cuda", data_a, N_SMALL, 1, True, constants=constants)
    fd = SMALFitter("cuda", data_b, N_SMALL, 1, True, constants=constants)
    ev = fc.stage_targets(*sets[1], torch.cuda.Stream())
    fc.swap_targets(ev)
    la_, _ = fc(list(range(N_SMALL)), STAGE1, 1)
    lb_, _ = fd(list(range(N_SMALL)), STAGE1, 1)
    assert float(la_) == float(lb_)


def test_cuda_path_matches_reference_smalfitter_golden(constants):
    """The CUDA path against vectors the UNMODIFIED reference `SMALFitter` produced (tests/golden/fitter_golden.npz, made by
    tests/golden/make_fitter_golden.py: its forward, get_temporal and torch autograd, with the PyTorch3D renderer replaced by
    a stand-in that renders with the oracle): initial parameter block, loss terms, temporal terms and every gradient, for
    three rows of the reference's OPT_WEIGHTS, one of them on a sub-window."""
    import os
    from smalify_b200.smal_fitter import SMALFitter
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fitter_golden.npz"))
    S, N = int(g["S"]), int(g["N"])
    sil = torch.from_numpy(np.unpackbits(g["sil"])[:N * S * S].reshape(N, 1, S, S).astype(np.float32))
    data = (torch.zeros(N, 3, S, S), sil, torch.from_numpy(g["joints"]), torch.from_numpy(g["vis"]))
    f = SMALFitter("cuda", data, N, 1, True, constants=constants)
    assert np.abs(f.betas.detach().cpu().numpy() - g["init_betas"]).max() <= 1e-7
    assert np.abs(f.log_beta_scales.detach().cpu().numpy().reshape(-1)[:6] - g["init_log_beta_scales"]).max() <= 1e-7
    assert np.abs(f.global_rotation.detach().cpu().numpy() - g["init_global_rotation"]).max() <= 1e-6
    assert float(f.joint_rotations.detach().abs().max()) == 0.0 and float(f.trans.detach().abs().max()) == 0.0
    names = ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")
    for stage in range(3):
        pre = "s%d_" % stage
        with torch.no_grad():
            for k in names:
                getattr(f, k).copy_(torch.from_numpy(g["p_" + k]).to(f.device).reshape(getattr(f, k).shape))
        for t in f.parameters():
            t.grad = None
            t.requires_grad_(True)
        w = [float(x) for x in g[pre + "weights"]]
        br = [int(i) for i in g[pre + "batch_range"]]
        loss, objs = f(br, w[:6], stage)
        jl, gl, tl = f.get_temporal(w[6])
        (loss + jl + gl + tl).backward()
        ref = float(g[pre + "loss"])
        assert abs(float(loss) - ref) <= 2e-5 * abs(ref) + 1e-6, (stage, float(loss), ref)
        for k in ("joint", "sil_reproj", "betas", "pose", "splay"):
            r = float(g[pre + "term_" + k])
            if not np.isnan(r):
                assert abs(float(objs[k]) - r) <= 3e-5 * abs(r) + 1e-6, (stage, k, float(objs[k]), r)
        for a, b in zip((jl, gl, tl), g[pre + "temporal"]):
            assert abs(float(a) - float(b)) <= 2e-5 * abs(float(b)) + 1e-9, (stage, float(a), float(b))
        for k in names:
            r = torch.from_numpy(g[pre + "grad_" + k])
            assert H.rel_err(getattr(f, k).grad.reshape(r.shape), r) < 1e-4, (stage, k)
