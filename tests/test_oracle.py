"""CPU: the oracle against the reference's golden vectors and its own consistency checks."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import smal_oracle as O
from smalify_b200 import constants as K

GOLD = os.path.join(os.path.dirname(__file__), "golden", "smal_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float64, 2e-6)])
def test_smal_matches_reference_golden(constants, gold, dtype, tol):
    """Golden vectors were produced by the unmodified reference SMAL (tests/golden/make_golden.py)."""
    m = O.OracleModel.from_constants(constants, dtype)
    T = lambda k: torch.from_numpy(gold[k]).to(dtype)  # noqa: E731
    betas, ls, gl, jo, tr = [T(k).requires_grad_(True) for k in ("betas", "logscale", "glob", "joint", "trans")]
    theta = torch.cat([gl[:, None], jo], 1)
    verts, joints, vs = O.smal_forward(m, betas, theta, ls)
    verts = verts + tr[:, None]
    joints = joints + tr[:, None]
    assert (verts.detach() - T("verts")).abs().max() < tol
    assert (joints.detach() - T("joints")).abs().max() < tol
    assert (vs.detach() - T("v_shaped")).abs().max() < tol
    probe = (verts * T("probe_v")).sum() + (joints * T("probe_j")).sum()
    grads = torch.autograd.grad(probe, [betas, ls, gl, jo, tr])
    for k, g in zip(("g_betas", "g_logscale", "g_glob", "g_joint", "g_trans"), grads):
        ref = T(k)
        assert (g - ref).abs().max() <= 5e-5 * max(1.0, float(ref.abs().max())), k


def test_pose_prior_matches_reference_golden(constants, gold):
    m = O.OracleModel.from_constants(constants, torch.float32)
    theta = torch.cat([torch.from_numpy(gold["glob"])[:, None], torch.from_numpy(gold["joint"])], 1)
    res = (((theta.reshape(-1, 105) - m.pose_mean) @ m.pose_prec) * m.pose_use) ** 2
    assert torch.allclose(res, torch.from_numpy(gold["pose_res"]), rtol=1e-5, atol=1e-4)


def test_global_rotation_init_is_head_on():
    R = O.rodrigues(torch.tensor([K.GLOBAL_ROT_INIT], dtype=torch.float64))[0]
    want = torch.tensor([[0.0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=torch.float64)   # SURVEY 8a footnote 1
    assert (R - want).abs().max() < 1e-7      # eps inside the norm (batch_lbs.py:39)


def test_projection_hand_computed():
    S = 256
    pts = torch.tensor([[[0.0, 0.0, 0.0], [0.3, -0.2, 0.5]]], dtype=torch.float64)
    rc = O.project_points_screen(pts, S)
    assert torch.allclose(rc[0, 0], torch.tensor([(S - 1) / 2.0, (S - 1) / 2.0], dtype=torch.float64))
    f = 1 / math.tan(math.radians(30))
    zv = 2.7 - 0.5
    want_col = (S - 1) / 2 * (1 - (-f * 0.3 / zv))
    want_row = (S - 1) / 2 * (1 - (f * -0.2 / zv))
    assert abs(float(rc[0, 1, 0]) - want_row) < 1e-9 and abs(float(rc[0, 1, 1]) - want_col) < 1e-9


def _quad_mesh(z=2.0, half=0.3):
    v = torch.tensor([[-half, -half, z], [half, -half, z], [half, half, z], [-half, half, z]], dtype=torch.float64)
    f = torch.tensor([[0, 1, 2], [0, 2, 3]])
    return v, f


def test_silhouette_hard_limit_and_orientation():
    """Deep inside a face alpha -> 1, far outside -> 0; the NDC x axis points left, y up."""
    S = 32
    v, f = _quad_mesh()
    v = v.clone()
    v[:, 0] += 0.4           # shift towards +x NDC = towards column 0
    a = O.soft_silhouette(v, f, S)
    cols = torch.nonzero(a.sum(0) > 0.5)[:, 0]
    assert cols.float().mean() < S / 2 - 3
    px = lambda i: 1 - (2 * i + 1) / S   # noqa: E731
    for r in range(S):
        for c in range(S):
            x, y = px(c), px(r)
            inside = (0.1 + 0.05 < x < 0.7 - 0.05) and (-0.3 + 0.05 < y < 0.3 - 0.05)
            outside = not ((0.1 - 0.04 < x < 0.7 + 0.04) and (-0.3 - 0.04 < y < 0.3 + 0.04))
            if inside:
                # the diagonal of the quad passes through the interior: near it alpha is 1-0.25
                assert a[r, c] > 0.7
            if outside:
                assert a[r, c] == 0.0


def test_k_cap_keeps_nearest():
    """150 stacked coplanar-in-xy triangles at increasing depth: only the 100 nearest count."""
    S = 8
    tris, faces = [], []
    n = 150
    for i in range(n):
        z = 1.0 + 0.01 * i
        # triangle whose edge passes 0.01 NDC from the pixel centre (0.125, 0.125): partial alpha each
        tris += [[0.135, -1.0, z], [0.135, 1.0, z], [1.5, 0.0, z]]
        faces.append([3 * i, 3 * i + 1, 3 * i + 2])
    v = torch.tensor(tris, dtype=torch.float64)
    f = torch.tensor(faces)
    a = O.soft_silhouette(v, f, S)
    r = c = 3                      # pixel centre x = y = 1 - 7/8 = 0.125
    d2 = 0.01 ** 2
    p = 1 / (1 + math.exp(d2 / O.SIGMA))
    want = 1 - (1 - p) ** 100
    assert abs(float(a[r, c]) - want) < 1e-12
    a_all = O.soft_silhouette(v, f, S, k_faces=1000)
    assert abs(float(a_all[r, c]) - (1 - (1 - p) ** 150)) < 1e-12


def test_raster_finite_differences(constants, oracle64):
    """fp64 central differences of sum(alpha * probe) w.r.t. a few vertex coordinates."""
    p = O.FitParams.initial(oracle64, 1, K.GLOBAL_ROT_INIT)
    theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
    verts, _, _ = O.smal_forward(oracle64, p.betas[None], theta, p.log_beta_scales[None])
    ndc = O.world_to_ndc(verts)[0].detach()
    S = 32
    probe = torch.randn(S, S, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    x = ndc.clone().requires_grad_(True)
    (O.soft_silhouette(x, oracle64.faces, S) * probe).sum().backward()
    g = x.grad
    idx = torch.argsort(g[:, :2].abs().sum(1), descending=True)[:4]
    h = 1e-7
    for vi in idx.tolist():
        for ci in (0, 1):
            xp, xm = ndc.clone(), ndc.clone()
            xp[vi, ci] += h
            xm[vi, ci] -= h
            fd = ((O.soft_silhouette(xp, oracle64.faces, S) * probe).sum() - (O.soft_silhouette(xm, oracle64.faces, S) * probe).sum()) / (2 * h)
            assert abs(float(fd) - float(g[vi, ci])) <= 2e-4 * max(1.0, abs(float(fd))), (vi, ci, float(fd), float(g[vi, ci]))
    assert float(g[:, 2].abs().max()) == 0.0      # z carries no gradient in the silhouette shader


def test_c_rasteriser_matches_torch_oracle(constants, oracle64):
    from oracle import raster_c
    p = O.FitParams.initial(oracle64, 1, K.GLOBAL_ROT_INIT)
    theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
    verts, _, _ = O.smal_forward(oracle64, p.betas[None], theta, p.log_beta_scales[None])
    ndc = O.world_to_ndc(verts)[0].detach()
    S = 48
    a_t, st = O.soft_silhouette(ndc, oracle64.faces, S, return_stats=True)
    for mode in (0, 1):
        a_c, _, st_c = raster_c.soft_silhouette_np(ndc.numpy(), constants.faces, S, mode)
        assert np.abs(a_c - a_t.numpy()).max() < 1e-5
        assert st_c["n_frag"] == st["n_frag"] and st_c["capped"] == st["capped"]
    x = ndc.clone().requires_grad_(True)
    ga = torch.randn(S, S, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    (O.soft_silhouette(x, oracle64.faces, S) * ga).sum().backward()
    _, gv, _ = raster_c.soft_silhouette_np(ndc.numpy(), constants.faces, S, 1, grad_alpha=ga.numpy())
    assert np.abs(gv - x.grad.numpy()).max() <= 1e-4 * float(x.grad.abs().max())


def test_temporal_and_losses_shapes(constants, oracle64):
    p = O.FitParams.initial(oracle64, 3, K.GLOBAL_ROT_INIT)
    p.trans = torch.tensor([[0.0, 0, 0], [0.1, 0, 0], [0.1, 0.2, 0]], dtype=torch.float64)
    jl, gl, tl = O.temporal_terms(p, 100.0)
    assert float(jl) == 0.0 and float(gl) == 0.0
    assert abs(float(tl) - 100.0 * (0.01 / 3 + 0.04 / 3)) < 1e-12


def test_joint_limit_term_matches_reference_golden(constants):
    """Row 8f-4: the limits table and the hinge of priors/joint_limits_prior.py (golden made by
    tests/golden/make_limits_golden.py from the unmodified reference) against constants.joint_limits() and
    the oracle's restatement of the commented-out term (smal_fitter.py:146-151)."""
    import numpy as np
    import torch
    from oracle import smal_oracle as O
    from smalify_b200 import constants as K
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "joint_limits_golden.npz"))
    lo, hi = K.joint_limits()
    assert lo.shape == hi.shape == (K.N_POSE, 3)
    assert np.array_equal(lo[:32].reshape(-1), g["min_values"].astype(np.float32))
    assert np.array_equal(hi[:32].reshape(-1), g["max_values"].astype(np.float32))
    assert np.all(np.isinf(lo[32:])) and np.all(np.isinf(hi[32:]))           # ears: unbounded
    # oracle term on the golden batch (ears at zero: no contribution)
    m = O.OracleModel.from_constants(constants, torch.float64)
    B = g["x"].shape[0]
    p = O.FitParams.initial(m, B, K.GLOBAL_ROT_INIT)
    q = torch.zeros(B, K.N_POSE, 3, dtype=torch.float64)
    q[:, :32] = torch.from_numpy(g["x"]).reshape(B, 32, 3)
    p.joint_rotations = q
    w = (0.0, 0.0, 0.0, 0.0, 7.0, 0.0)
    sil = torch.zeros(B, 1, 16, 16)
    total, objs = O.fitter_forward(m, p, sil, torch.zeros(B, 25, 2), torch.zeros(B, 25), range(B), w, 16,
                                   joint_limits=(lo.astype(np.float64), hi.astype(np.float64)))
    want = 7.0 * g["hinge"].sum() / (B * K.N_POSE * 3)       # torch.mean over (B, 34, 3)
    lo32, hi32 = g["min_values"], g["max_values"]
    # (float32 table vs the reference's float64 literals: 1e-7 relative)
    assert abs(float(objs["limit"]) - want) < 1e-6 * abs(want)
    assert set(objs) == {"limit"}
    # without limits the weight is ignored, as in the reference
    total0, objs0 = O.fitter_forward(m, p, sil, torch.zeros(B, 25, 2), torch.zeros(B, 25), range(B), w, 16)
    assert objs0 == {} and float(total0) == 0.0


@pytest.mark.parametrize("dtype,ltol,gtol", [(torch.float32, 2e-6, 1e-4), (torch.float64, 2e-5, 1e-4)])
def test_loss_assembly_matches_reference_smalfitter_golden(constants, dtype, ltol, gtol):
    """tests/golden/fitter_golden.npz holds what the UNMODIFIED reference `SMALFitter` computed (forward, get_temporal,
    torch autograd; its PyTorch3D renderer replaced by a stand-in that renders with this oracle -- see
    tests/golden/make_fitter_golden.py) for three rows of its own OPT_WEIGHTS: initial parameter block, every loss term,
    the temporal terms and all gradients."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fitter_golden.npz"))
    S, N = int(g["S"]), int(g["N"])
    m = O.OracleModel.from_constants(constants, dtype)
    init = O.FitParams.initial(m, N, K.GLOBAL_ROT_INIT)
    if dtype == torch.float32:
        assert np.array_equal(init.betas.numpy(), g["init_betas"]) and np.array_equal(init.log_beta_scales.numpy(), g["init_log_beta_scales"])
    assert np.abs(init.global_rotation.numpy() - g["init_global_rotation"]).max() < 1e-6
    sil = torch.from_numpy(np.unpackbits(g["sil"])[:N * S * S].reshape(N, 1, S, S).astype(np.float32))
    joints, vis = torch.from_numpy(g["joints"]), torch.from_numpy(g["vis"])
    names = ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")
    for stage in range(3):
        pre = "s%d_" % stage
        p = O.FitParams(**{k: torch.from_numpy(g["p_" + k]).to(dtype).requires_grad_(True) for k in names})
        w = g[pre + "weights"]
        br = [int(i) for i in g[pre + "batch_range"]]
        loss, objs = O.fitter_forward(m, p, sil, joints, vis, br, w[:6], S)
        jl, gl, tl = O.temporal_terms(p, float(w[6]))
        (loss + jl + gl + tl).backward()
        assert abs(float(loss) - float(g[pre + "loss"])) <= ltol * abs(float(g[pre + "loss"])), stage
        for k in ("joint", "sil_reproj", "betas", "pose", "splay"):
            ref = float(g[pre + "term_" + k])
            assert (k in objs) == (not math.isnan(ref)), (stage, k)
            if k in objs:
                assert abs(float(objs[k]) - ref) <= ltol * max(abs(ref), 1e-12), (stage, k, float(objs[k]), ref)
        for a, b in zip((jl, gl, tl), g[pre + "temporal"]):
            assert abs(float(a) - float(b)) <= ltol * max(abs(float(b)), 1e-12), stage
        for k in names:
            ref = torch.from_numpy(g[pre + "grad_" + k]).double()
            got = getattr(p, k).grad.double()
            assert float((got - ref).abs().max()) <= gtol * max(float(ref.abs().max()), 1e-12), (stage, k)
