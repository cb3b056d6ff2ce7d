"""The restated reference CPU path (torch-CPU SMAL + C restatement of the PyTorch3D CPU
rasteriser + restated losses + torch Adam), used by bench.py as ``cpu_baseline`` and
``--impl reference``.  CPU ORACLE: never imported by the product."""
from __future__ import annotations

import time

import torch

from . import raster_c
from . import smal_oracle as O


def c_silhouette_fn(mode: int):
    """mode 0: the reference's every-face-every-pixel loop, single thread;
    mode 1: OpenMP + per-row culling on all host cores."""
    def fn(m, verts, S):
        ndc = O.world_to_ndc(verts)
        return raster_c.SoftSilhouetteC.apply(ndc, m.faces, S, mode)[:, None]
    return fn


def time_cpu_epochs(constants, data, window, weights, w_temp, lr, image_size, n_steps: int, mode: int = 1,
                    warmup: int = 0):
    """Runs ``warmup + n_steps`` epochs (forward over all frames of ``data`` + temporal + backward + Adam
    step, optimize_to_joints.py:117-137) on the host in float32, as the reference does, and returns
    (seconds per epoch, final loss)."""
    m = O.OracleModel.from_constants(constants, torch.float32)
    rgb, sil, joints, vis = data
    n = sil.shape[0]
    from smalify_b200.constants import GLOBAL_ROT_INIT
    p = O.FitParams.initial(m, n, GLOBAL_ROT_INIT)
    for t in p.tensors():
        t.requires_grad_(True)
    opt = torch.optim.Adam(p.tensors(), lr=lr, betas=(0.5, 0.999))
    fn = c_silhouette_fn(mode)
    loss = None
    t0 = None
    for it in range(warmup + n_steps):
        if it == warmup:
            t0 = time.perf_counter()
        opt.zero_grad()
        loss = O.epoch_loss(m, p, sil, joints, vis, window, weights, w_temp, image_size, silhouette_fn=fn)
        loss.backward()
        opt.step()
    dt = (time.perf_counter() - t0) / max(n_steps, 1)
    return dt, float(loss)
