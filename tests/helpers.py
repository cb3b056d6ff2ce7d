"""Shared builders for the parity tests (oracle side = checker only)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import smal_oracle as O
from smalify_b200 import constants as K
from smalify_b200 import synthetic


def oracle_renderer(m: O.OracleModel, S: int):
    def render(gt):
        n = gt["global_rotation"].shape[0]
        dt = m.dtype
        theta = torch.cat([gt["global_rotation"][:, None], gt["joint_rotations"]], dim=1).to(dt)
        verts, joints, _ = O.smal_forward(m, gt["betas"].to(dt).expand(n, 20), theta, gt["log_beta_scales"].to(dt).expand(n, 6))
        verts = verts + gt["trans"].to(dt)[:, None]
        joints = joints + gt["trans"].to(dt)[:, None]
        alpha = O.render_silhouettes(m, verts, S)[:, 0]
        kp = O.project_points_screen(joints[:, list(O.CANONICAL)], S)
        return (alpha > 0.5).to(torch.uint8), kp.float()
    return render


def perturbed_params(m: O.OracleModel, gt: dict, seed: int, scale: float = 1.0) -> O.FitParams:
    """A parameter state away from both the init and the ground truth."""
    g = torch.Generator().manual_seed(seed)
    n = gt["global_rotation"].shape[0]
    dt = m.dtype
    return O.FitParams(
        global_rotation=(gt["global_rotation"] + 0.1 * scale * torch.randn(n, 3, generator=g)).to(dt),
        joint_rotations=(gt["joint_rotations"] + 0.1 * scale * torch.randn(n, 34, 3, generator=g)).to(dt),
        betas=(gt["betas"] + 0.2 * scale * torch.randn(20, generator=g)).to(dt),
        log_beta_scales=(gt["log_beta_scales"] + 0.05 * scale * torch.randn(6, generator=g)).to(dt),
        trans=(gt["trans"] + 0.03 * scale * torch.randn(n, 3, generator=g)).to(dt))


def oracle_loss_and_grads(m, p: O.FitParams, data, batch_range, weights, S, w_temp=None, joint_limits=None, focal=None):
    """focal: optional 0-d tensor; its gradient is returned under the key 'focal'."""
    rgb, sil, joints, vis = data
    for t in p.tensors():
        t.requires_grad_(True)
        t.grad = None
    if focal is not None:
        focal.requires_grad_(True)
        focal.grad = None
    loss, objs = O.fitter_forward(m, p, sil, joints, vis, batch_range, weights, S, joint_limits=joint_limits, focal=focal)
    if w_temp is not None:
        jl, gl, tl = O.temporal_terms(p, w_temp)
        loss = loss + jl + gl + tl
    loss.backward()
    grads = {k: (getattr(p, k).grad.clone() if getattr(p, k).grad is not None else torch.zeros_like(getattr(p, k)))
             for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")}
    for t in p.tensors():
        t.requires_grad_(False)
    if focal is not None:
        grads["focal"] = focal.grad.clone()
        focal.requires_grad_(False)
    return float(loss), {k: float(v) for k, v in objs.items()}, grads


def load_params_into(fitter, p: O.FitParams):
    with torch.no_grad():
        for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans"):
            getattr(fitter, k).copy_(getattr(p, k).detach().float().to(fitter.device))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (the SURVEY 8d gradient metric)."""
    b = b.double().cpu()
    a = a.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def bench_subsequence(constants, m: O.OracleModel, idx, S: int, n_total: int = 128, seed: int = 0):
    """Frames `idx` of the synthetic `n_total`-frame sequence bench.py fits (same ground truth, keypoint noise and
    visibility rows), with targets rendered by the oracle.  Returns (data_batch, gt_params_of_the_subset)."""
    return synthetic.make_subsequence(constants, n_total, idx, S, oracle_renderer(m, S), seed=seed)


def record_result(name: str, payload: dict):
    """Parity figures the GPU tests measure (pixel-flip counts, error maxima) are appended to
    gpurun_out/parity_results.json so that they travel back from the GPU box."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_results.json")
        cur = {}
        if os.path.exists(path):
            with open(path) as fh:
                cur = json.load(fh)
        cur[name] = payload
        with open(path, "w") as fh:
            json.dump(cur, fh, indent=1, sort_keys=True)
    except OSError:
        pass
