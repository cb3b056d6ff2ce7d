#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference, imported in place.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports ``smal_model.smal_torch.SMAL`` and ``priors.pose_prior_35.Prior`` from
the SMALify checkout (a chumpy stub is registered so the pickles load), runs
them on seeded inputs in float32 exactly as ``SMALFitter.forward`` calls them
(smal_fitter.py:122-127,153-157) and stores inputs, outputs and torch-autograd
gradients of fixed scalar probes in ``tests/golden/smal_golden.npz``.

The rasteriser half of the path lives in PyTorch3D 0.2.5, which is not
available here, so there are no goldens for it (see oracle/smal_oracle.py).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SMALIFY_REF", "/root/reference")


def import_reference():
    sys.path.insert(0, REPO)
    from smalify_b200.model_io import _ChStub
    for name in ("chumpy", "chumpy.ch"):
        mod = types.ModuleType(name)
        mod.Ch = _ChStub
        sys.modules[name] = mod
    os.chdir(REF)                              # config.py uses relative paths
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "smal_fitter"))
    from smal_model.smal_torch import SMAL
    from priors.pose_prior_35 import Prior
    import config
    return SMAL, Prior, config


def main():
    SMAL, Prior, config = import_reference()
    torch.manual_seed(0)
    smal = SMAL("cpu", shape_family_id=1)
    prior = Prior(config.WALKING_PRIOR_FILE, "cpu")
    unity = np.load(config.UNITY_SHAPE_PRIOR)
    mean = torch.from_numpy(unity["mean"][:-1]).float()

    B = 3
    g = torch.Generator().manual_seed(1234)
    betas = (mean[:20] + 0.3 * torch.randn(20, generator=g)).expand(B, 20).clone().requires_grad_(True)
    logscale = (mean[20:] + 0.1 * torch.randn(6, generator=g)).expand(B, 6).clone().requires_grad_(True)
    init = torch.tensor([-1.2091996, -1.2091996, -1.2091996])
    glob = (init[None] + 0.2 * torch.randn(B, 3, generator=g)).requires_grad_(True)
    joint = (0.25 * torch.randn(B, 34, 3, generator=g))
    joint[0] = 0.0                                   # frame 0: the all-zero init pose (Rodrigues at 0)
    joint = joint.requires_grad_(True)
    trans = (0.1 * torch.randn(B, 3, generator=g)).requires_grad_(True)

    theta = torch.cat([glob[:, None], joint], dim=1)
    verts, joints, Rs, v_shaped = smal(betas, theta, betas_logscale=logscale)
    verts = verts + trans[:, None]
    joints = joints + trans[:, None]

    # fixed linear probes -> reference gradients through the whole body model
    pv = torch.randn(verts.shape, generator=g)
    pj = torch.randn(joints.shape, generator=g)
    probe = (verts * pv).sum() + (joints * pj).sum()
    grads = torch.autograd.grad(probe, [betas, logscale, glob, joint, trans])

    pose_res = prior(theta)                            # (B,105) squared residuals
    out = dict(
        betas=betas.detach().numpy(), logscale=logscale.detach().numpy(), glob=glob.detach().numpy(),
        joint=joint.detach().numpy(), trans=trans.detach().numpy(),
        verts=verts.detach().numpy(), joints=joints.detach().numpy(), v_shaped=v_shaped.detach().numpy(),
        Rs=Rs.detach().numpy(), probe_v=pv.numpy(), probe_j=pj.numpy(),
        g_betas=grads[0].numpy(), g_logscale=grads[1].numpy(), g_glob=grads[2].numpy(),
        g_joint=grads[3].numpy(), g_trans=grads[4].numpy(),
        pose_res=pose_res.detach().numpy(),
    )
    path = os.path.join(HERE, "smal_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
