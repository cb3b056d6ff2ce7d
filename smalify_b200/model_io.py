"""Load the SMAL model + priors and build the sparse tables the kernels use.

Two sources give the same ``SmalConstants``:

* ``load_from_smalify_data(data_root, shape_family)`` reads the pickles a
  SMALify checkout ships (``data/SMALST/smpl_models/*.pkl``, ``data/priors/*``)
  and applies the same preprocessing as ``SMAL.__init__``
  (smal_model/smal_torch.py:24-96) and ``SMALFitter.__init__``
  (smal_fitter/smal_fitter.py:43-74).
* ``load_asset(path)`` reads the compact ``.npz`` written by
  ``tools/export_assets.py`` (used on boxes that have no SMALify checkout).

All floating point constants are float32, exactly as the reference holds them
after ``torch.Tensor(...)``.
"""
from __future__ import annotations

import io
import os
import pickle
from dataclasses import dataclass, field

import numpy as np

from . import constants as C

_ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def default_asset_path(shape_family: int = 1) -> str:
    return os.path.join(_ASSET_DIR, f"smal_family{shape_family}.npz")


# --------------------------------------------------------------------------
# un-pickling without chumpy: the pickles reference chumpy.ch.Ch objects whose
# state dict carries the ndarray under 'x'.
# --------------------------------------------------------------------------
class _ChStub:
    def __setstate__(self, state):
        self.__dict__.update(state)

    @property
    def r(self):
        return np.asarray(self.__dict__["x"])

    @property
    def shape(self):
        return self.r.shape


class _Unpickler(pickle._Unpickler):  # pure-python unpickler: find_class hook
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChStub
        return super().find_class(module, name)


def _load_pickle(path: str):
    with open(path, "rb") as f:
        u = _Unpickler(io.BytesIO(f.read()))
        u.encoding = "latin1"
        return u.load()


def _dense(x):
    if isinstance(x, _ChStub):
        return x.r
    if hasattr(x, "todense"):
        return np.asarray(x.todense())
    return np.asarray(x)


# --------------------------------------------------------------------------
@dataclass
class SmalConstants:
    """Everything constant during a fit. Dense fp32 copies follow the reference;
    the sparse tables are derived from them (``build_tables``)."""

    shape_family: int
    v_template: np.ndarray      # (V,3) f32, family mean baked in + symmetrised
    shapedirs: np.ndarray       # (20, V*3) f32, interleaved xyz
    j_regressor: np.ndarray     # (V,35) f32 dense
    weights: np.ndarray         # (V,35) f32 dense
    faces: np.ndarray           # (F,3) int32
    parents: np.ndarray         # (35,) int32, root = -1
    # priors
    unity_mean: np.ndarray      # (26,) f32
    unity_prec: np.ndarray      # (26,26) f32
    cluster_mean: np.ndarray    # (20,) f32
    cluster_prec: np.ndarray    # (20,20) f32
    pose_mean: np.ndarray       # (105,) f32
    pose_prec: np.ndarray       # (105,105) f32  ('pic')
    pose_use: np.ndarray        # (105,) f32 mask (first 3 = 0)
    # optional: BADJA rs_dog visibility rows (201,25) for synthetic inputs
    badja_visibility: np.ndarray | None = None
    tables: dict = field(default_factory=dict)

    def __post_init__(self):
        if not self.tables:
            self.tables = build_tables(self)


def _symmetry_axis_indices() -> np.ndarray:
    out = []
    for a, b in C.SYMMETRY_AXIS_RUNS:
        out.extend(range(a, b + 1))
    return np.asarray(out, dtype=np.int64)


def _symmetrise_template(v: np.ndarray, sym_idx: np.ndarray) -> np.ndarray:
    """Behaviour of align_smal_template_to_symmetry_axis (smal_basics.py:7-37):
    centre on the scalar mean of all coordinates, put the mid-line vertices on
    y=0 and mirror the left half onto the right through symIdx."""
    v = np.array(v, dtype=np.float64)
    axis = _symmetry_axis_indices()
    v = v - v.mean()
    v[:, 1] -= v[axis, 1].mean()
    v[axis, 1] = 0.0
    left = v[:, 1] < 0
    v[left[sym_idx]] = v[left] * np.array([1.0, -1.0, 1.0])
    right = v[:, 1] > 0
    if int(left.sum()) != int(right.sum()):
        raise ValueError("SMAL template is not left/right balanced after symmetrisation")
    return v


def _prec_from_cov(cov: np.ndarray) -> np.ndarray:
    # smal_fitter.py:54-55 / :65-66
    inv = np.linalg.inv(cov + 1e-5 * np.eye(cov.shape[0]))
    return np.linalg.cholesky(inv)


def load_from_smalify_data(data_root: str, shape_family: int = 1,
                           pose_prior: str = "walking_toy_symmetric_pose_prior_with_cov_35parts.pkl",
                           badja_json: str | None = None) -> SmalConstants:
    """``data_root`` is the SMALify ``data`` directory (config.py:7)."""
    mdir = os.path.join(data_root, "SMALST", "smpl_models")
    dd = _load_pickle(os.path.join(mdir, "my_smpl_00781_4_all.pkl"))
    data = _load_pickle(os.path.join(mdir, "my_smpl_data_00781_4_all.pkl"))
    sym_idx = np.asarray(_load_pickle(os.path.join(mdir, "symIdx.pkl")), dtype=np.int64)

    faces = np.asarray(dd["f"]).astype(np.int32)
    v_template = np.asarray(dd["v_template"], dtype=np.float64)
    nv = v_template.shape[0]
    shp = _dense(dd["shapedirs"])                       # (V,3,41)
    n_all = shp.shape[-1]
    shapedir64 = shp.reshape(-1, n_all).T.copy()        # (41, V*3)  smal_torch.py:53-54

    if shape_family != -1:
        betas = np.asarray(data["cluster_means"][shape_family], dtype=np.float64)
        v_template = v_template + (betas[None, :] @ shapedir64).reshape(nv, 3)   # smal_torch.py:66-69
    v_sym = _symmetrise_template(v_template, sym_idx)

    posedirs = _dense(dd["posedirs"])
    if np.abs(posedirs).max() != 0.0:
        raise ValueError("posedirs is not identically zero: the fused kernels skip the pose "
                         "blendshape (smal_torch.py:138-142) and cannot run this model")

    parents = np.asarray(dd["kintree_table"][0]).astype(np.int64)
    parents[0] = -1

    # priors ---------------------------------------------------------------
    pri_dir = os.path.join(data_root, "priors")
    unity = np.load(os.path.join(pri_dir, "unity_betas.npz"))
    unity_mean = np.asarray(unity["mean"][:-1], dtype=np.float32)            # smal_fitter.py:50-52
    unity_prec = _prec_from_cov(np.asarray(unity["cov"])[:-1, :-1])
    fam = max(shape_family, 0)
    ccov = np.asarray(data["cluster_cov"])[fam]
    cluster_prec = _prec_from_cov(ccov).astype(np.float32)[:C.N_BETAS, :C.N_BETAS]   # smal_fitter.py:68
    cluster_mean = np.asarray(data["cluster_means"][fam], dtype=np.float32)[:C.N_BETAS]

    pp = _load_pickle(os.path.join(pri_dir, pose_prior))
    pose_prec = _dense(pp["pic"]).astype(np.float32)
    pose_mean = np.asarray(pp["mean_pose"]).astype(np.float32)
    pose_use = np.ones(105, dtype=np.float32)
    pose_use[:3] = 0.0                                                       # pose_prior_35.py:78-81

    vis = None
    if badja_json is None:
        cand = os.path.join(data_root, "BADJA", "joint_annotations", "rs_dog.json")
        badja_json = cand if os.path.exists(cand) else None
    if badja_json is not None:
        vis = load_badja_visibility(badja_json)

    return SmalConstants(
        shape_family=shape_family,
        v_template=v_sym.astype(np.float32),
        shapedirs=shapedir64[:C.N_BETAS].astype(np.float32),
        j_regressor=_dense(dd["J_regressor"]).T.astype(np.float32).copy(),
        weights=_dense(dd["weights"]).astype(np.float32),
        faces=faces,
        parents=parents.astype(np.int32),
        unity_mean=unity_mean,
        unity_prec=unity_prec.astype(np.float32),
        cluster_mean=cluster_mean,
        cluster_prec=cluster_prec,
        pose_mean=pose_mean,
        pose_prec=pose_prec,
        pose_use=pose_use,
        badja_visibility=vis,
    )


def load_badja_visibility(json_path: str) -> np.ndarray:
    """Visibility rows mapped to the 25 keypoints as load_badja_sequence does
    (data_loader.py:41-42,64-66): unannotated classes are invisible."""
    import json
    with open(json_path) as f:
        ann = json.load(f)
    cls = np.asarray(C.BADJA_ANNOTATED_CLASSES)
    rows = []
    for a in ann:
        v = np.asarray(a["visibility"], dtype=bool)[cls]   # -1 indexes the last entry, then masked
        v = v & (cls != -1)
        rows.append(v)
    return np.stack(rows).astype(np.uint8)


_ASSET_KEYS = ("v_template", "shapedirs", "faces", "parents", "unity_mean", "unity_prec",
               "cluster_mean", "cluster_prec", "pose_mean", "pose_prec", "pose_use")


def save_asset(c: SmalConstants, path: str) -> None:
    """Compact form: dense regressor / skinning matrices stored as COO."""
    jr, jc = np.nonzero(c.j_regressor)
    wr, wc = np.nonzero(c.weights)
    arrays = {k: getattr(c, k) for k in _ASSET_KEYS}
    arrays.update(
        shape_family=np.int32(c.shape_family),
        jreg_rows=jr.astype(np.int32), jreg_cols=jc.astype(np.int32), jreg_vals=c.j_regressor[jr, jc],
        w_rows=wr.astype(np.int32), w_cols=wc.astype(np.int32), w_vals=c.weights[wr, wc],
    )
    if c.badja_visibility is not None:
        arrays["badja_visibility"] = c.badja_visibility
    np.savez_compressed(path, **arrays)


def load_asset(path: str | None = None, shape_family: int = 1) -> SmalConstants:
    path = path or default_asset_path(shape_family)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"SMAL asset {path} not found; run tools/export_assets.py against a SMALify checkout "
            "or use load_from_smalify_data()")
    z = np.load(path)
    nv = z["v_template"].shape[0]
    jreg = np.zeros((nv, C.N_JOINTS), dtype=np.float32)
    jreg[z["jreg_rows"], z["jreg_cols"]] = z["jreg_vals"]
    w = np.zeros((nv, C.N_JOINTS), dtype=np.float32)
    w[z["w_rows"], z["w_cols"]] = z["w_vals"]
    return SmalConstants(
        shape_family=int(z["shape_family"]),
        j_regressor=jreg, weights=w,
        badja_visibility=z["badja_visibility"] if "badja_visibility" in z.files else None,
        **{k: z[k] for k in _ASSET_KEYS},
    )


# --------------------------------------------------------------------------
# derived tables
# --------------------------------------------------------------------------
MAX_INFLUENCES = 8     # max non-zero skinning weights per vertex (measured: 8)


def scale_axis_table() -> np.ndarray:
    """(35,3) int32: which log-scale entry scales axis a of joint j, -1 = none
    (batch_lbs.py:107-121)."""
    t = -np.ones((C.N_JOINTS, 3), dtype=np.int32)
    for a, b, idx in C.SCALE_GROUPS:
        t[a:b + 1] = np.asarray(idx, dtype=np.int32)
    return t


def joint_levels(parents: np.ndarray):
    depth = np.zeros(len(parents), dtype=np.int32)
    for j in range(1, len(parents)):
        depth[j] = depth[parents[j]] + 1      # parents precede children in SMAL ordering
    order = np.argsort(depth, kind="stable").astype(np.int32)
    nlev = int(depth.max()) + 1
    level_start = np.zeros(nlev + 1, dtype=np.int32)
    for d in depth:
        level_start[d + 1] += 1
    level_start = np.cumsum(level_start).astype(np.int32)
    return depth, order, level_start


def build_tables(c: SmalConstants) -> dict:
    nv = c.v_template.shape[0]
    nf = c.faces.shape[0]
    assert np.all(c.parents[1:] < np.arange(1, C.N_JOINTS)), "parents must precede children"
    t: dict = {}

    # skinning weights, ELL by vertex (joint id, weight), padded with weight 0
    infl_j = np.zeros((nv, MAX_INFLUENCES), dtype=np.int32)
    infl_w = np.zeros((nv, MAX_INFLUENCES), dtype=np.float32)
    cnt = np.zeros(nv, dtype=np.int32)
    rows, cols = np.nonzero(c.weights)
    for r, j in zip(rows, cols):
        k = cnt[r]
        if k >= MAX_INFLUENCES:
            raise ValueError("vertex with more than 8 skinning influences")
        infl_j[r, k] = j
        infl_w[r, k] = c.weights[r, j]
        cnt[r] += 1
    t["skin_joint"] = infl_j
    t["skin_weight"] = infl_w
    t["skin_count"] = cnt

    # skinning weights, CSC by joint (vertex, weight) for the joint-centric backward
    order = np.lexsort((rows, cols))
    t["skinT_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=C.N_JOINTS))]).astype(np.int32)
    t["skinT_vert"] = rows[order].astype(np.int32)
    t["skinT_weight"] = c.weights[rows[order], cols[order]].astype(np.float32)

    # joint regressor by joint (vertex, weight) for the 35 regressed joints, then the
    # 6 picked vertices as one-entry rows -> 41 model joints
    jr, jc = np.nonzero(c.j_regressor)
    o = np.lexsort((jr, jc))
    ptr = [0]
    verts, vals = [], []
    counts = np.bincount(jc, minlength=C.N_JOINTS)
    pos = 0
    for j in range(C.N_JOINTS):
        sel = o[pos:pos + counts[j]]
        verts.extend(jr[sel].tolist())
        vals.extend(c.j_regressor[jr[sel], j].tolist())
        pos += counts[j]
        ptr.append(len(verts))
    t["jreg_ptr"] = np.asarray(ptr, dtype=np.int32)            # (36,) regressed joints only
    t["jreg_vert"] = np.asarray(verts, dtype=np.int32)
    t["jreg_weight"] = np.asarray(vals, dtype=np.float32)
    for pv in C.PICKED_VERTS:
        verts.append(pv)
        vals.append(1.0)
        ptr.append(len(verts))
    t["mj_ptr"] = np.asarray(ptr, dtype=np.int32)              # (42,) all 41 model joints
    t["mj_vert"] = np.asarray(verts, dtype=np.int32)
    t["mj_weight"] = np.asarray(vals, dtype=np.float32)

    # transposed: by vertex, list of (model joint, weight); most vertices have none
    mjv = np.asarray(verts, dtype=np.int64)
    mjj = np.repeat(np.arange(C.N_MODEL_JOINTS), np.diff(ptr))
    o2 = np.lexsort((mjj, mjv))
    t["mjT_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(mjv, minlength=nv))]).astype(np.int32)
    t["mjT_joint"] = mjj[o2].astype(np.int32)
    t["mjT_weight"] = np.asarray(vals, dtype=np.float32)[o2]
    # same for the regressed joints only (shape path: J = Jreg^T v_shaped)
    o3 = np.lexsort((jc, jr))
    t["jregT_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(jr, minlength=nv))]).astype(np.int32)
    t["jregT_joint"] = jc[o3].astype(np.int32)
    t["jregT_weight"] = c.j_regressor[jr[o3], jc[o3]].astype(np.float32)

    # vertex -> incident (face, corner) CSR, for the deterministic face->vertex gather
    fv = c.faces.reshape(-1).astype(np.int64)
    fidx = np.repeat(np.arange(nf), 3) * 4 + np.tile(np.arange(3), nf)       # face*4 + corner
    o4 = np.lexsort((fidx, fv))
    t["v2f_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(fv, minlength=nv))]).astype(np.int32)
    t["v2f_fc"] = fidx[o4].astype(np.int32)

    depth, jorder, lstart = joint_levels(c.parents)
    t["joint_depth"] = depth
    t["joint_order"] = jorder
    t["level_start"] = lstart
    t["scale_axis"] = scale_axis_table()
    t["keypoint_joint"] = np.asarray(C.CANONICAL_MODEL_JOINTS, dtype=np.int32)
    return t
