"""CPU restatement of the reference's colour renderer (test infrastructure; row 8f-3).

smal_fitter/p3d_renderer.py:41-59,70-72 builds MeshRenderer(MeshRasterizer(blur_radius=0,
faces_per_pixel=1), HardPhongShader(lights=PointLights(location=[[0, 0, 3]]))) with constant vertex
colours.  PyTorch3D 0.2.5 is not available here (see smal_oracle.py): **parity unpinned** -- this follows
the published 0.2.5 algorithm (rasterize_meshes naive path, shading.phong_shading, lighting.diffuse /
specular, blending.hard_rgb_blend, Meshes.verts_normals_packed) with its defaults: light ambient /
diffuse / specular 0.5 / 0.3 / 0.2, materials 1, shininess 64, white background, barycentrics not
perspective-corrected.  float64 numpy, one frame at a time, every face against its own pixel box.
"""
from __future__ import annotations

import numpy as np

from . import smal_oracle as O


def vertex_normals(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Meshes.verts_normals_packed: the faces' (v1 - v0) x (v2 - v0) accumulated on their vertices, normalised."""
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    fn = np.cross(v1 - v0, v2 - v0)
    n = np.zeros_like(verts)
    for k in range(3):
        np.add.at(n, faces[:, k], fn)
    return n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-6)


def render_color(verts: np.ndarray, faces: np.ndarray, S: int, color, focal=None):
    """(V,3) world vertices -> (3,S,S) image and the (S,S) face-index map (-1 = background)."""
    import torch
    verts = np.asarray(verts, np.float64)
    ndc = O.world_to_ndc(torch.from_numpy(verts), focal).numpy()
    normals = vertex_normals(verts, faces)
    zbuf = np.full((S, S), np.inf)
    fidx = np.full((S, S), -1, np.int64)
    bary = np.zeros((S, S, 3))
    xs = 1.0 - (2.0 * np.arange(S) + 1.0) / S           # pixel centres; +X left, +Y up
    for f, (i0, i1, i2) in enumerate(faces):
        (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = ndc[i0], ndc[i1], ndc[i2]
        area = (x2 - x0) * (y1 - y0) - (y2 - y0) * (x1 - x0)
        if max(z0, z1, z2) < 0 or abs(area) <= O.K_EPS:
            continue
        cols = np.nonzero((xs >= min(x0, x1, x2)) & (xs <= max(x0, x1, x2)))[0]
        rows = np.nonzero((xs >= min(y0, y1, y2)) & (xs <= max(y0, y1, y2)))[0]
        if len(cols) == 0 or len(rows) == 0:
            continue
        px, py = np.meshgrid(xs[cols], xs[rows])
        den = area + O.K_EPS
        w0 = ((px - x1) * (y2 - y1) - (py - y1) * (x2 - x1)) / den
        w1 = ((px - x2) * (y0 - y2) - (py - y2) * (x0 - x2)) / den
        w2 = ((px - x0) * (y1 - y0) - (py - y0) * (x1 - x0)) / den
        pz = w0 * z0 + w1 * z1 + w2 * z2
        hit = (w0 > 0) & (w1 > 0) & (w2 > 0) & (pz >= 0)
        sub = zbuf[np.ix_(rows, cols)]
        better = hit & (pz < sub)                         # strict: the lower face id wins a tie
        if not better.any():
            continue
        rr, cc = np.nonzero(better)
        zbuf[rows[rr], cols[cc]] = pz[rr, cc]
        fidx[rows[rr], cols[cc]] = f
        bary[rows[rr], cols[cc]] = np.stack([w0[rr, cc], w1[rr, cc], w2[rr, cc]], -1)
    img = np.ones((S, S, 3))
    m = fidx >= 0
    fv = faces[fidx[m]]
    w = bary[m][:, :, None]
    P = (verts[fv] * w).sum(1)
    N = (normals[fv] * w).sum(1)
    N = N / np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-6)
    D = np.array([0.0, 0.0, 3.0]) - P
    D = D / np.maximum(np.linalg.norm(D, axis=1, keepdims=True), 1e-6)
    cosang = (N * D).sum(1)
    diffuse = 0.3 * np.maximum(cosang, 0.0)
    Vd = np.array([0.0, 0.0, O.CAMERA_DISTANCE]) - P
    Vd = Vd / np.maximum(np.linalg.norm(Vd, axis=1, keepdims=True), 1e-6)
    Rf = -D + 2.0 * cosang[:, None] * N
    alpha = np.maximum((Vd * Rf).sum(1), 0.0) * (cosang > 0)
    specular = 0.2 * alpha ** 64
    img[m] = (0.5 + diffuse)[:, None] * np.asarray(color, np.float64)[None] + specular[:, None]
    return np.transpose(img, (2, 0, 1)), fidx
