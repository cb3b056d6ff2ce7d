"""Fit-quality metrics (the reference defines none)."""
from __future__ import annotations

import torch


def keypoint_l2(proj: torch.Tensor, target: torch.Tensor, visibility: torch.Tensor) -> float:
    """Mean over visible joints of ||proj - target||_2 in pixels; (row, col) layout."""
    v = visibility.bool().to(proj.device)
    d = torch.linalg.norm(proj.double() - target.to(proj.device).double(), dim=-1)
    return float(d[v].mean()) if bool(v.any()) else 0.0


def silhouette_iou(alpha: torch.Tensor, target: torch.Tensor) -> float:
    """count({alpha>0.5} & T) / count({alpha>0.5} | T), averaged over frames."""
    n = alpha.shape[0]
    a = alpha.reshape(n, -1) > 0.5
    t = target.to(alpha.device).reshape(n, -1) > 0.5
    inter = (a & t).sum(1).double()
    union = (a | t).sum(1).double().clamp(min=1)
    return float((inter / union).mean())
