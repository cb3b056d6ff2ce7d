"""Seeded synthetic sequences: BADJA frames and masks are not shipped
with the reference, so benchmarks and parity tests fit to a synthetic animal.

``ground_truth_params`` only draws parameters (pure torch, CPU).  Turning them into
targets needs a renderer; callers pass one:  ``render(params) -> (sil u8 (n,S,S),
keypoints f32 (n,25,2))``.  ``gpu_renderer`` uses libsmalfit itself (bench / GPU
tests), the CPU tests pass the oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import constants as K


def _rotvec_compose_y(phi: float, base_rotvec) -> np.ndarray:
    """rot-vec of R_y(phi) @ R(base)."""
    from scipy.spatial.transform import Rotation as R
    return (R.from_euler("y", phi) * R.from_rotvec(np.asarray(base_rotvec))).as_rotvec()


def ground_truth_params(constants, n: int, seed: int = 0, per_frame_shapes: bool = False) -> dict:
    g = torch.Generator().manual_seed(seed)
    mean = torch.from_numpy(np.asarray(constants.unity_mean)).float()
    n_shapes = n if per_frame_shapes else 1
    betas = mean[:20][None] + 0.3 * torch.randn(n_shapes, 20, generator=g)
    lbs = mean[20:26][None].repeat(n_shapes, 1)
    t = np.arange(n)
    phi = 0.5 * np.sin(2 * math.pi * t / 64.0) + math.pi / 2
    glob = np.stack([_rotvec_compose_y(float(p), K.GLOBAL_ROT_INIT) for p in phi]).astype(np.float32)
    direction = torch.randn(34, 3, generator=g)
    phase = 2 * math.pi * torch.rand(34, generator=g)
    tt = torch.from_numpy(t).float()
    joint = 0.25 * torch.sin(2 * math.pi * tt[:, None] / 32.0 + phase[None])[:, :, None] * direction[None]
    joint = joint.clamp(-0.6, 0.6)
    trans = torch.stack([0.1 * torch.sin(2 * math.pi * tt / 128.0), torch.zeros(n), torch.full((n,), 0.9)], dim=1)
    return dict(betas=betas if per_frame_shapes else betas[0], log_beta_scales=lbs if per_frame_shapes else lbs[0],
                global_rotation=torch.from_numpy(glob), joint_rotations=joint.float(), trans=trans.float())


def make_subsequence(constants, n_total: int, idx, image_size: int, render, seed: int = 0, kp_noise_px: float = 1.5,
                     per_frame_shapes: bool = False, pad_to=None):
    """Frames `idx` of the seeded `n_total`-frame sequence: the same ground truth, keypoint noise and visibility rows
    whichever subset is asked for (a rank of a sharded run renders its own frames only).  Returns (data_batch,
    gt_params_of_the_subset) with data_batch = (rgb, sil (n,1,S,S) f32, joints (n,25,2), visibility (n,25)) in the
    layout of the reference loaders.  pad_to=(offset, length): the arrays have `length` frames with the subset at
    [offset, offset + n) and zeros elsewhere (a fitter that keeps the sequence's full parameter layout but only
    ever evaluates its own shard)."""
    gt_all = ground_truth_params(constants, n_total, seed, per_frame_shapes)
    idx = list(idx)
    n = len(idx)
    shared = () if per_frame_shapes else ("betas", "log_beta_scales")
    gt = {k: (v if k in shared else v[idx]) for k, v in gt_all.items()}
    sil_u8, kp = render(gt)
    g = torch.Generator().manual_seed(seed + 1)
    noise = kp_noise_px * torch.randn(n_total, K.N_KEYPOINTS, 2, generator=g)
    joints = kp.cpu().float() + noise[idx]
    if constants.badja_visibility is not None:
        rows = np.asarray(constants.badja_visibility)
        vis = torch.from_numpy(rows[np.asarray(idx) % len(rows)].astype(np.float32))
    else:
        vis = torch.ones(n, K.N_KEYPOINTS)
    sil = sil_u8.cpu().float().reshape(n, 1, image_size, image_size)
    if pad_to is not None:
        off, length = int(pad_to[0]), int(pad_to[1])
        full = (torch.zeros(length, 1, image_size, image_size), torch.zeros(length, K.N_KEYPOINTS, 2), torch.zeros(length, K.N_KEYPOINTS))
        for dst, src in zip(full, (sil, joints, vis)):
            dst[off:off + n] = src
        sil, joints, vis = full
        n = length
    # (rgb is only looked at by the visualisation: a 1x1 placeholder broadcast over the frames keeps big runs small)
    rgb = torch.zeros(1, 3, 1, 1).expand(n, 3, image_size, image_size)
    return (rgb, sil, joints, vis), gt


def make_sequence(constants, n: int, image_size: int, render, seed: int = 0, kp_noise_px: float = 1.5):
    """The whole seeded n-frame sequence: (data_batch, gt_params)."""
    return make_subsequence(constants, n, range(n), image_size, render, seed, kp_noise_px)


def gpu_renderer(constants, image_size: int, device="cuda", per_frame_shapes: bool = False):
    """render(params) through libsmalfit (targets are inputs, not results)."""
    from .smal_fitter import SMALFitter

    def render(gt):
        n = gt["global_rotation"].shape[0]
        blank = (None, torch.zeros(n, 1, image_size, image_size), torch.zeros(n, K.N_KEYPOINTS, 2),
                 torch.zeros(n, K.N_KEYPOINTS))
        f = SMALFitter(device, blank, n, constants.shape_family, True, constants=constants, per_frame_shapes=per_frame_shapes)
        with torch.no_grad():
            for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans"):
                getattr(f, k).copy_(gt[k].to(f.device))
            alpha, kp = f.render()
        return (alpha.reshape(n, image_size, image_size) > 0.5).to(torch.uint8).cpu(), kp.cpu()

    return render
