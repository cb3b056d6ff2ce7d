"""Visualisation of a fit (row 8f-3): the collage smal_fitter.py:209-272 writes every
``VIS_FREQUENCY`` epochs -- target keypoints | Phong rendering | overlay | silhouette agreement |
rendering turned by 180 degrees -- with the reference's joint markers (draw_smal_joints.py:9-46).
The colour pass runs in libsmalfit (``smalfit_render_color``); the rest is host-side image assembly.
"""
from __future__ import annotations

import numpy as np
import torch

from . import constants as K

# config.py:105-129 (marker per annotated keypoint); cv2 marker ids: CROSS 0, STAR 2, TRIANGLE_DOWN 6
_CROSS, _STAR, _TRI = 0, 2, 6
MARKER_TYPE = ([_TRI, _STAR, _CROSS] * 4 + [_CROSS, _TRI] + [_CROSS, _CROSS] + [_CROSS, _STAR] + [_TRI, _TRI]
               + [_CROSS, _CROSS] + [_CROSS, _CROSS] + [_STAR])
MARKER_COLORS = ([[230, 25, 75]] * 3 + [[255, 255, 25]] * 3 + [[60, 180, 75]] * 3 + [[0, 130, 200]] * 3
                 + [[240, 50, 230]] * 2 + [[255, 153, 204], [29, 98, 115]] + [[245, 130, 48]] * 2
                 + [[255, 153, 204], [29, 98, 115]] + [[0, 0, 0]] * 2 + [[128, 0, 0]] * 2 + [[240, 50, 230]])
MESH_COLOR = (0.0, 172.0 / 255.0, 223.0 / 255.0)         # config.py:60


def draw_joints(image: torch.Tensor, landmarks: torch.Tensor, visible=None) -> torch.Tensor:
    """(B,3,H,W) float [0,1] images with a marker per (row, col) landmark; invisible joints are parked along the
    top edge, 10 px apart (draw_smal_joints.py:33-40)."""
    import cv2
    img = np.ascontiguousarray((np.transpose(image.detach().cpu().numpy(), (0, 2, 3, 1)) * 255.0).astype(np.uint8))
    lm = landmarks.detach().cpu().numpy()
    vis = np.ones(lm.shape[:2], bool) if visible is None else visible.detach().cpu().numpy().astype(bool)
    out = []
    for im, pts, vs in zip(img, lm, vis):
        im = im.copy()
        parked = 0
        for j, ((y, x), v) in enumerate(zip(pts, vs)):
            if not v:
                x, y = parked * 10, 0
                parked += 1
            cv2.drawMarker(im, (int(x), int(y)), tuple(int(c) for c in MARKER_COLORS[j]), MARKER_TYPE[j], 8, thickness=3)
        out.append(im)
    return torch.from_numpy(np.transpose(np.stack(out, 0) / 255.0, (0, 3, 1, 2)).astype(np.float32))


@torch.no_grad()
def collage(fitter, batch_range=None) -> torch.Tensor:
    """(B, 3, S, 5 S) float [0,1]: the five panels of smal_fitter.py:254-262 for the fitter's current parameters."""
    br = list(batch_range) if batch_range is not None else list(range(fitter.num_images))
    dev = fitter.device
    verts = fitter.vertices(br)                                        # (B,V,3), trans included
    sil, joints2d = fitter.render(br)
    joints3d = fitter.model_joints(br)[:, list(K.CANONICAL_MODEL_JOINTS)]
    rendered = fitter.render_color(verts)
    # the same mesh seen from behind: centred on the vertex mean and turned by 180 degrees about y (:240-246)
    centre = verts.mean(dim=1, keepdim=True)
    flip = torch.tensor([-1.0, 1.0, -1.0], device=dev)
    rev = fitter.render_color((verts - centre) * flip)
    rev_joints = fitter.project_points((joints3d - centre) * flip)
    rgb = fitter.rgb_imgs[br].to(dev).float()
    target_sil = fitter.sil_imgs[br].to(dev).float()
    vis = fitter.target_visibility[br]
    overlay = rendered * 0.8 + rgb * 0.2
    sil_err = (1.0 - (target_sil - sil).abs()).expand_as(rgb).cpu()
    return torch.cat([
        draw_joints(rgb, fitter.target_joints[br], vis), draw_joints(rendered, joints2d, vis),
        draw_joints(overlay, joints2d, vis), sil_err, draw_joints(rev, rev_joints, vis)], dim=3)
