"""CPU ORACLE for the SMALify fitting path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product
(``smalify_b200``) never does.

What it is: a plain-torch (CPU, autograd, float32 or float64) restatement of the
reference hot path, written to follow the reference's own procedure step by
step (dense matmuls, the 34-step kinematic loop with explicit inverses, a
K-slot per-pixel fragment table and ``prod`` blend) so that it is an
*independent* check of the closed forms the CUDA kernels use.

Pinning status
--------------
* SMAL body model, pose prior, shape prior: PINNED against the reference's own
  ``smal_model.smal_torch.SMAL`` / ``priors.pose_prior_35.Prior`` imported in
  the build container; golden vectors in ``tests/golden/smal_golden.npz``
  (generator: ``tests/golden/make_golden.py``).
* Camera, soft rasteriser, blend, raster backward: **parity unpinned**.  They
  live in PyTorch3D 0.2.5 (``requirements.txt:60``), which is not vendored in
  the reference and not installable here (no network).  The restatement
  follows PyTorch3D 0.2.5's published algorithm
  (renderer/cameras.py, renderer/mesh/rasterize_meshes.py,
  csrc/rasterize_meshes/rasterize_meshes_cpu.cpp, csrc/utils/geometry_utils.h,
  renderer/blending.py::sigmoid_alpha_blend) as called from
  ``smal_fitter/p3d_renderer.py:22-39,61-68``; self-consistency is checked by
  fp64 finite differences, hard-coverage limits and hand-computed projections
  in ``tests/test_oracle.py``.
* Loss terms, temporal term, parameter block: restated from
  ``smal_fitter/smal_fitter.py:107-190`` and PINNED against the unmodified
  ``SMALFitter`` itself, imported in the build container with its PyTorch3D
  ``Renderer`` replaced by a stand-in that renders with this module
  (``tests/test_fitter_vs_reference.py``: losses bit-identical, gradients 3e-6);
  the loader's tables for every shape family in ``tests/test_loader_vs_reference.py``.
  Stage loop: restated from ``smal_fitter/optimize_to_joints.py:90-137`` and PINNED
  the same way: the reference's ``main()`` is run and every optimiser step it takes
  (trainable set, Adam settings and state, gradients) is checked against this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# Constants restated from the reference (see smalify_b200/constants.py for the
# citations); duplicated here on purpose so the oracle stands alone.
CAMERA_DISTANCE = 2.7                     # p3d_renderer.py:22
FOV_DEG = 60.0                            # OpenGLPerspectiveCameras default
SIGMA = 1e-4                              # p3d_renderer.py:26
BLUR_RADIUS = math.log(1.0 / 1e-4 - 1.0) * SIGMA   # p3d_renderer.py:29
K_FACES = 100                             # p3d_renderer.py:30
K_EPS = 1e-8                              # PyTorch3D kEpsilon
PICKED = (1863, 26, 2124, 150, 3055, 1097)          # smal_torch.py:176-183
CANONICAL = (10, 9, 8, 20, 19, 18, 14, 13, 12, 24, 23, 22, 25, 31, 33, 34,
             35, 36, 38, 37, 39, 40, 15, 15, 28)    # config.py:77-88
TORSO = (2, 5, 8, 11, 12, 23)                       # config.py:75


@dataclass
class OracleModel:
    """Dense model tensors in a chosen dtype (what SMAL.__init__ keeps)."""
    v_template: torch.Tensor    # (V,3)
    shapedirs: torch.Tensor     # (20, V*3)
    j_regressor: torch.Tensor   # (V,35)
    weights: torch.Tensor       # (V,35)
    faces: torch.Tensor         # (F,3) int64
    parents: np.ndarray         # (35,)
    pose_mean: torch.Tensor
    pose_prec: torch.Tensor
    pose_use: torch.Tensor
    shape_mean: torch.Tensor    # 26 (unity) or 20
    shape_prec: torch.Tensor
    use_unity_prior: bool
    dtype: torch.dtype

    @staticmethod
    def from_constants(c, dtype=torch.float64, use_unity_prior=True) -> "OracleModel":
        t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)  # noqa: E731
        return OracleModel(
            v_template=t(c.v_template), shapedirs=t(c.shapedirs), j_regressor=t(c.j_regressor),
            weights=t(c.weights), faces=torch.from_numpy(np.asarray(c.faces).astype(np.int64)),
            parents=np.asarray(c.parents).astype(np.int64),
            pose_mean=t(c.pose_mean), pose_prec=t(c.pose_prec), pose_use=t(c.pose_use),
            shape_mean=t(c.unity_mean if use_unity_prior else c.cluster_mean),
            shape_prec=t(c.unity_prec if use_unity_prior else c.cluster_prec),
            use_unity_prior=use_unity_prior, dtype=dtype)


# --------------------------------------------------------------------------
# SMAL body model (smal_model/smal_torch.py:99-189, batch_lbs.py)
# --------------------------------------------------------------------------
def rodrigues(theta: torch.Tensor) -> torch.Tensor:
    """batch_lbs.py:33-52.  theta (N,3) -> (N,3,3).  eps is added to every
    component before the norm."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)       # (N,1)
    r = theta / angle
    c = torch.cos(angle)[:, :, None]
    s = torch.sin(angle)[:, :, None]
    outer = r[:, :, None] * r[:, None, :]
    zero = torch.zeros_like(r[:, 0])
    skew = torch.stack([
        torch.stack([zero, -r[:, 2], r[:, 1]], dim=1),
        torch.stack([r[:, 2], zero, -r[:, 0]], dim=1),
        torch.stack([-r[:, 1], r[:, 0], zero], dim=1)], dim=1)        # batch_lbs.py:9-31
    eye = torch.eye(3, dtype=theta.dtype)[None]
    return c * eye + (1 - c) * outer + s * skew


def scale_mask(dtype) -> torch.Tensor:
    """(6,105) mask of batch_lbs.py:107-124."""
    m = torch.zeros(35, 3, 6, dtype=dtype)
    legs = list(range(7, 11)) + list(range(11, 15)) + list(range(17, 21)) + list(range(21, 25))
    tail = list(range(25, 32))
    ears = [33, 34]
    m[legs, 2, 0] = 1.0
    m[legs, 0, 1] = 1.0
    m[legs, 1, 1] = 1.0
    m[tail, 0, 2] = 1.0
    m[tail, 1, 3] = 1.0
    m[tail, 2, 3] = 1.0
    m[ears, 1, 4] = 1.0
    m[ears, 2, 5] = 1.0
    return m.reshape(105, 6).t().contiguous()


def global_rigid_transformation(Rs, Js, parents, logscale):
    """batch_lbs.py:75-170 followed literally: per joint S_parent^-1 R S, chained
    4x4 products, then A = G - pad(G [J;0])."""
    n = Rs.shape[0]
    dt = Rs.dtype
    scaling = torch.exp(logscale @ scale_mask(dt)).reshape(n, 35, 3)
    S = torch.diag_embed(scaling)

    def make_a(R, t):
        top = torch.cat([R, t[:, :, None]], dim=2)                    # (n,3,4)
        bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=dt).expand(n, 1, 4)
        return torch.cat([top, bottom], dim=1)

    results = [make_a(Rs[:, 0], Js[:, 0])]
    for i in range(1, 35):
        p = int(parents[i])
        rot = torch.inverse(S[:, p]) @ Rs[:, i] @ S[:, i]
        results.append(results[p] @ make_a(rot, Js[:, i] - Js[:, p]))
    G = torch.stack(results, dim=1)                                   # (n,35,4,4)
    new_j = G[:, :, :3, 3]
    j_h = torch.cat([Js, torch.zeros(n, 35, 1, dtype=dt)], dim=2)[..., None]
    init_bone = G @ j_h                                               # (n,35,4,1)
    A = G - torch.nn.functional.pad(init_bone, (3, 0))
    return new_j, A


def smal_forward(m: OracleModel, betas, theta, logscale):
    """betas (B,20), theta (B,35,3), logscale (B,6) -> verts (B,V,3), joints (B,41,3).
    posedirs is identically zero in the shipped model and is dropped
    (smal_torch.py:138-142 adds zeros)."""
    B = betas.shape[0]
    V = m.v_template.shape[0]
    v_shaped = m.v_template[None] + (betas @ m.shapedirs).reshape(B, V, 3)     # :115
    J = torch.einsum("bvc,vj->bjc", v_shaped, m.j_regressor)                    # :125-128
    Rs = rodrigues(theta.reshape(-1, 3)).reshape(B, 35, 3, 3)                   # :135
    _, A = global_rigid_transformation(Rs, J, m.parents, logscale)              # :145
    T = (m.weights[None] @ A.reshape(B, 35, 16)).reshape(B, V, 4, 4)            # :152-158
    v_h = torch.cat([v_shaped, torch.ones(B, V, 1, dtype=betas.dtype)], dim=2)
    verts = (T @ v_h[..., None])[:, :, :3, 0]                                   # :159-163
    joints = torch.einsum("bvc,vj->bjc", verts, m.j_regressor)                  # :171-174
    joints = torch.cat([joints, verts[:, list(PICKED)]], dim=1)                 # :176-184
    return verts, joints, v_shaped


# --------------------------------------------------------------------------
# Camera (PyTorch3D 0.2.5 look_at_view_transform(2.7,0,0) + OpenGLPerspective)
# --------------------------------------------------------------------------
def world_to_ndc(verts: torch.Tensor, focal=None) -> torch.Tensor:
    """(…,3) world -> (x_ndc, y_ndc, z_view).  R = diag(-1,1,-1), T = (0,0,2.7);
    x_ndc = f x_view / z_view with f = 1/tan(30 deg); the rasteriser's z is view z
    (MeshRasterizer.transform).  focal (extension): replaces f."""
    f = 1.0 / math.tan(math.radians(FOV_DEG) / 2.0) if focal is None else focal
    xv = -verts[..., 0]
    yv = verts[..., 1]
    zv = CAMERA_DISTANCE - verts[..., 2]
    return torch.stack([f * xv / zv, f * yv / zv, zv], dim=-1)


def project_points_screen(points: torch.Tensor, image_size: int, focal=None) -> torch.Tensor:
    """cameras.transform_points_screen(points, (S,S))[:, :, [1, 0]]
    (p3d_renderer.py:67-68): returns (row, col)."""
    ndc = world_to_ndc(points, focal)
    col = (image_size - 1.0) / 2.0 * (1.0 - ndc[..., 0])
    row = (image_size - 1.0) / 2.0 * (1.0 - ndc[..., 1])
    return torch.stack([row, col], dim=-1)


# --------------------------------------------------------------------------
# Soft rasteriser + sigmoid blend (PyTorch3D 0.2.5 semantics, see module doc)
# --------------------------------------------------------------------------
def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist2(px, py, ax, ay, bx, by):
    """PointLineDistanceForward: squared distance to segment a-b, degenerate
    segment -> distance to b."""
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    safe = torch.where(l2 <= K_EPS, torch.ones_like(l2), l2)
    t = ((bax * (px - ax) + bay * (py - ay)) / safe).clamp(0.0, 1.0)
    qx = ax + t * bax - px
    qy = ay + t * bay - py
    d = qx * qx + qy * qy
    dd = (px - bx) ** 2 + (py - by) ** 2
    return torch.where(l2 <= K_EPS, dd, d)


def candidate_pairs(vn: np.ndarray, faces: np.ndarray, S: int):
    """Conservative list of (face, row, col) whose pixel centre may lie in the
    blur-expanded bbox of the face.  Pure index generation (numpy)."""
    r = math.sqrt(BLUR_RADIUS)
    fx = vn[faces, 0]
    fy = vn[faces, 1]
    xmin, xmax = fx.min(1) - r, fx.max(1) + r
    ymin, ymax = fy.min(1) - r, fy.max(1) + r
    # pixel centre x(c) = 1 - (2c+1)/S  =>  c = ((1-x) S - 1)/2
    c_lo = np.clip(np.floor(((1.0 - xmax) * S - 1.0) / 2.0).astype(np.int64) - 1, 0, S - 1)
    c_hi = np.clip(np.ceil(((1.0 - xmin) * S - 1.0) / 2.0).astype(np.int64) + 1, 0, S - 1)
    r_lo = np.clip(np.floor(((1.0 - ymax) * S - 1.0) / 2.0).astype(np.int64) - 1, 0, S - 1)
    r_hi = np.clip(np.ceil(((1.0 - ymin) * S - 1.0) / 2.0).astype(np.int64) + 1, 0, S - 1)
    off = (xmax < -1.0 - 2.0 / S) | (xmin > 1.0 + 2.0 / S) | (ymax < -1.0 - 2.0 / S) | (ymin > 1.0 + 2.0 / S)
    nc = np.where(off, 0, c_hi - c_lo + 1)
    nr = np.where(off, 0, r_hi - r_lo + 1)
    cnt = nc * nr
    total = int(cnt.sum())
    f_idx = np.repeat(np.arange(len(faces)), cnt)
    start = np.repeat(np.cumsum(cnt) - cnt, cnt)
    local = np.arange(total) - start
    ncr = np.repeat(np.maximum(nc, 1), cnt)
    rows = np.repeat(r_lo, cnt) + local // ncr
    cols = np.repeat(c_lo, cnt) + local % ncr
    return f_idx, rows, cols


def soft_silhouette(verts_ndc: torch.Tensor, faces: torch.Tensor, S: int,
                    k_faces: int = K_FACES, return_stats: bool = False):
    """One mesh.  verts_ndc (V,3) [x_ndc, y_ndc, z_view] -> alpha (S,S).

    Per pixel, per face (RasterizeMeshesNaiveCpu / CheckPixelInsideFace):
    skip if zmax<0, |area|<=eps, pixel outside bbox +- sqrt(blur); barycentrics
    from edge functions over (area + eps); pz = sum w z; skip pz<0; d2 = min
    squared segment distance; inside = all w>0; skip if !inside and d2>=blur;
    keep the K smallest pz (ties: lower face index).  Blend
    (sigmoid_alpha_blend): alpha = 1 - prod_k (1 - sigmoid(-signed_k/sigma)).
    """
    dt = verts_ndc.dtype
    fnp = faces.numpy()
    f_idx, rows, cols = candidate_pairs(verts_ndc.detach().numpy().astype(np.float64), fnp, S)
    f_t = torch.from_numpy(f_idx)
    tri = verts_ndc[faces[f_t]]                      # (P,3,3)
    x0, y0, z0 = tri[:, 0, 0], tri[:, 0, 1], tri[:, 0, 2]
    x1, y1, z1 = tri[:, 1, 0], tri[:, 1, 1], tri[:, 1, 2]
    x2, y2, z2 = tri[:, 2, 0], tri[:, 2, 1], tri[:, 2, 2]
    px = 1.0 - (2.0 * torch.from_numpy(cols).to(dt) + 1.0) / S
    py = 1.0 - (2.0 * torch.from_numpy(rows).to(dt) + 1.0) / S
    rad = math.sqrt(BLUR_RADIUS)
    xmin = torch.minimum(torch.minimum(x0, x1), x2)
    xmax = torch.maximum(torch.maximum(x0, x1), x2)
    ymin = torch.minimum(torch.minimum(y0, y1), y2)
    ymax = torch.maximum(torch.maximum(y0, y1), y2)
    zmax = torch.maximum(torch.maximum(z0, z1), z2)
    out_bbox = (px > xmax + rad) | (px < xmin - rad) | (py > ymax + rad) | (py < ymin - rad)
    area = _edge(x2, y2, x0, y0, x1, y1)
    ok = (~out_bbox) & (zmax >= 0) & ~((area <= K_EPS) & (area >= -K_EPS))
    den = area + K_EPS
    den = torch.where(ok, den, torch.ones_like(den))
    w0 = _edge(px, py, x1, y1, x2, y2) / den
    w1 = _edge(px, py, x2, y2, x0, y0) / den
    w2 = _edge(px, py, x0, y0, x1, y1) / den
    pz = w0 * z0 + w1 * z1 + w2 * z2
    ok = ok & (pz >= 0)
    d2 = torch.minimum(torch.minimum(_seg_dist2(px, py, x0, y0, x1, y1),
                                     _seg_dist2(px, py, x0, y0, x2, y2)),
                       _seg_dist2(px, py, x1, y1, x2, y2))
    inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
    ok = ok & (inside | (d2 < BLUR_RADIUS))
    signed = torch.where(inside, -d2, d2)

    sel = torch.nonzero(ok)[:, 0]
    pix = torch.from_numpy(rows * S + cols)[sel]
    pz_s = pz.detach()[sel].double().numpy()
    f_s = f_idx[sel.numpy()]
    pix_np = pix.numpy()
    order = np.lexsort((f_s, pz_s, pix_np))          # by pixel, then pz, then face index
    pix_o = pix_np[order]
    first = np.r_[True, pix_o[1:] != pix_o[:-1]]
    seg_start = np.maximum.accumulate(np.where(first, np.arange(len(pix_o)), 0))
    rank = np.arange(len(pix_o)) - seg_start
    keep = rank < k_faces
    upix, inv = np.unique(pix_o, return_inverse=True)
    slot_pix = torch.from_numpy(inv[keep])
    slot_k = torch.from_numpy(rank[keep])
    src = sel[torch.from_numpy(order[keep])]
    one_minus = torch.ones(len(upix), k_faces, dtype=dt)
    prob = torch.sigmoid(-signed[src] / SIGMA)
    one_minus = one_minus.index_put((slot_pix, slot_k), 1.0 - prob)
    alpha_t = 1.0 - torch.prod(one_minus, dim=1)
    alpha = torch.zeros(S * S, dtype=dt).index_put((torch.from_numpy(upix),), alpha_t).reshape(S, S)
    if return_stats:
        counts = np.bincount(inv, minlength=len(upix))
        stats = dict(n_pair=int((~out_bbox).sum()), n_frag=int(ok.sum()),
                     touched=int(len(upix)), capped=int((counts > k_faces).sum()),
                     max_frag=int(counts.max()) if len(counts) else 0)
        return alpha, stats
    return alpha


def render_silhouettes(m: OracleModel, verts: torch.Tensor, S: int, focal=None) -> torch.Tensor:
    """(B,V,3) world verts -> (B,1,S,S) like Renderer.forward's first output."""
    ndc = world_to_ndc(verts, focal)
    return torch.stack([soft_silhouette(ndc[b], m.faces, S) for b in range(verts.shape[0])])[:, None]


# --------------------------------------------------------------------------
# SMALFitter.forward / get_temporal (smal_fitter.py:107-190)
# --------------------------------------------------------------------------
@dataclass
class FitParams:
    global_rotation: torch.Tensor    # (N,3)
    joint_rotations: torch.Tensor    # (N,34,3)
    betas: torch.Tensor              # (20,)
    log_beta_scales: torch.Tensor    # (6,)
    trans: torch.Tensor              # (N,3)

    def tensors(self):
        return [self.global_rotation, self.joint_rotations, self.betas, self.log_beta_scales, self.trans]

    @staticmethod
    def initial(m: OracleModel, n: int, global_init) -> "FitParams":
        dt = m.dtype
        return FitParams(
            global_rotation=torch.tensor(global_init, dtype=dt).repeat(n, 1),
            joint_rotations=torch.zeros(n, 34, 3, dtype=dt),
            betas=m.shape_mean[:20].clone(),
            log_beta_scales=(m.shape_mean[20:26].clone() if m.use_unity_prior else torch.zeros(6, dtype=dt)),
            trans=torch.zeros(n, 3, dtype=dt))


def fitter_forward(m: OracleModel, p: FitParams, sil, target_joints, visibility, batch_range, weights,
                   image_size: int, return_aux: bool = False, silhouette_fn=None, joint_limits=None, focal=None):
    """SMALFitter.forward.  sil (N,1,S,S), target_joints (N,25,2) (row,col),
    visibility (N,25) {0,1}.  weights = (w_j2d, w_reproj, w_betas, w_pose, w_limit, w_splay)."""
    w_j2d, w_reproj, w_betas, w_pose, w_limit, w_splay = [float(w) for w in weights]
    br = list(batch_range)
    B = len(br)
    g = p.global_rotation[br]
    q = p.joint_rotations[br]
    betas = p.betas.expand(B, 20)
    ls = p.log_beta_scales.expand(B, 6)
    theta = torch.cat([g[:, None], q], dim=1)
    verts, joints, _ = smal_forward(m, betas, theta, ls)
    verts = verts + p.trans[br][:, None]
    joints = joints + p.trans[br][:, None]
    kp3d = joints[:, list(CANONICAL)]
    proj = project_points_screen(kp3d, image_size, focal)
    objs = {}
    aux = {}
    if w_j2d > 0:
        vis = visibility[br].bool()
        rj = torch.where(vis[..., None], proj, torch.full_like(proj, -1.0))
        tj = torch.where(vis[..., None], target_joints[br].to(m.dtype), torch.full_like(proj, -1.0))
        objs["joint"] = w_j2d * torch.mean((rj - tj) ** 2)                       # :140-144
    if w_pose > 0:
        res = ((theta.reshape(B, 105) - m.pose_mean[None]) @ m.pose_prec) * m.pose_use
        objs["pose"] = w_pose * torch.mean(res ** 2)                             # :153-157
    if w_limit > 0 and joint_limits is not None:
        # the term the reference keeps commented out (:146-151); joint_limits = (min, max), each (34, 3)
        lo, hi = (torch.as_tensor(a, dtype=m.dtype) for a in joint_limits)
        zeros = torch.zeros_like(q)
        objs["limit"] = w_limit * torch.mean(torch.max(q - hi, zeros) + torch.max(lo - q, zeros))
    if w_splay > 0:
        objs["splay"] = w_splay * torch.sum(q[:, :, [0, 2]] ** 2)                # :159-160
    if w_betas > 0:
        allb = torch.cat([betas, ls], dim=1) if m.use_unity_prior else betas
        res = (allb - m.shape_mean[None]) @ m.shape_prec
        objs["betas"] = w_betas * torch.mean(res ** 2)                           # :162-171
    if w_reproj > 0 or return_aux:
        sil_r = silhouette_fn(m, verts, image_size) if silhouette_fn else render_silhouettes(m, verts, image_size, focal)
        aux["silhouettes"] = sil_r
        if w_reproj > 0:
            objs["sil_reproj"] = w_reproj * torch.mean(torch.abs(sil_r - sil[br].to(m.dtype)))   # :172-173
    total = sum(objs.values()) if objs else torch.zeros((), dtype=m.dtype)
    if return_aux:
        aux.update(proj=proj, verts=verts, joints=joints)
        return total, objs, aux
    return total, objs


def temporal_terms(p: FitParams, w_temp: float):
    """get_temporal (smal_fitter.py:177-190): returns (joint, global, trans)."""
    g, q, t = p.global_rotation, p.joint_rotations, p.trans
    n = g.shape[0]
    zero = torch.zeros((), dtype=g.dtype)
    if n < 2:
        return zero, zero.clone(), zero.clone()
    gl = (((g[:-1] - g[1:]) ** 2).mean(dim=1) * w_temp).sum()
    jl = (((q[:-1] - q[1:]) ** 2).reshape(n - 1, -1).mean(dim=1) * w_temp).sum()
    tl = (((t[:-1] - t[1:]) ** 2).mean(dim=1) * w_temp).sum()
    return jl, gl, tl


def epoch_loss(m, p, sil, tj, vis, window, weights, w_temp, image_size, silhouette_fn=None):
    """One epoch of optimize_to_joints.py:117-135: sum of window losses + temporal."""
    n = p.global_rotation.shape[0]
    acc = torch.zeros((), dtype=m.dtype)
    for j in range(0, n, window):
        loss, _ = fitter_forward(m, p, sil, tj, vis, range(j, min(n, j + window)), weights, image_size,
                                 silhouette_fn=silhouette_fn)
        acc = acc + loss
    jl, gl, tl = temporal_terms(p, w_temp)
    return acc + jl + gl + tl


def stage_visibility(visibility: torch.Tensor, stage_id: int) -> torch.Tensor:
    """optimize_to_joints.py:98-110."""
    if stage_id == 0:
        v = torch.zeros_like(visibility)
        v[:, list(TORSO)] = visibility[:, list(TORSO)]
        return v
    return visibility.clone()


def fit(m, p: FitParams, sil, tj, vis, window, schedule, image_size, allow_limb_scaling=True,
        iters_override=None, callback=None, silhouette_fn=None):
    """The stage loop of optimize_to_joints.py:90-137 with torch.optim.Adam."""
    for stage_id, row in enumerate(schedule):
        weights, w_temp, iters, lr = row[:6], row[6], int(row[7]), row[8]
        if iters_override is not None:
            iters = iters_override[stage_id]
        for t in p.tensors():
            t.requires_grad_(True)
            t.grad = None
        if stage_id == 0:
            p.joint_rotations.requires_grad_(False)
            p.betas.requires_grad_(False)
            p.log_beta_scales.requires_grad_(False)
        elif not (allow_limb_scaling and m.use_unity_prior):
            p.log_beta_scales.requires_grad_(False)
        opt = torch.optim.Adam(p.tensors(), lr=lr, betas=(0.5, 0.999))
        v = stage_visibility(vis, stage_id)
        for it in range(iters):
            opt.zero_grad()
            loss = epoch_loss(m, p, sil, tj, v, window, weights, w_temp, image_size, silhouette_fn=silhouette_fn)
            loss.backward()
            opt.step()
            if callback is not None:
                callback(stage_id, it, float(loss))
    for t in p.tensors():
        t.requires_grad_(False)
    return p


# --------------------------------------------------------------------------
# Quality metrics (the reference defines none)
# --------------------------------------------------------------------------
def keypoint_l2(proj, target, visibility) -> float:
    v = visibility.bool()
    d = torch.linalg.norm(proj.double() - target.double(), dim=-1)
    return float(d[v].mean()) if v.any() else 0.0


def silhouette_iou(alpha, target) -> float:
    a = alpha.reshape(alpha.shape[0], -1) > 0.5
    t = target.reshape(target.shape[0], -1) > 0.5
    inter = (a & t).sum(1).double()
    union = (a | t).sum(1).double().clamp(min=1)
    return float((inter / union).mean())
