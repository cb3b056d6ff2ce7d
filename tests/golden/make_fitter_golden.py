#!/usr/bin/env python
"""Golden vectors of the loss assembly from the UNMODIFIED reference `SMALFitter`, imported in place.

Run in the build container only (needs /root/reference):

    python tests/golden/make_fitter_golden.py

Constructs the reference's `SMALFitter` (smal_fitter/smal_fitter.py) on a small seeded sequence and, for the first three
rows of its own `config.OPT_WEIGHTS`, stores `forward`'s loss terms, `get_temporal`'s terms and the torch-autograd gradients
of their sum in `tests/golden/fitter_golden.npz`, together with every input.  PyTorch3D does not exist here, so the
reference's `p3d_renderer.Renderer` is replaced by a stand-in that renders with the oracle's restated camera and
rasteriser: the goldens pin parameter block, masks, SMAL, priors, loss terms, normalisers and the temporal term to the
reference -- not the rasteriser itself (tests/test_fitter_vs_reference.py is the same comparison run live).
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
S, N = 64, 3


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import helpers as H
    from oracle import smal_oracle as O
    from smalify_b200 import constants as K, model_io, synthetic
    from smalify_b200.model_io import _ChStub

    state = {}

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod

    stub("chumpy", Ch=_ChStub)
    stub("chumpy.ch", Ch=_ChStub)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("draw_smal_joints", SMALJointDrawer=type("SMALJointDrawer", (), {}))
    stub("utils", eul_to_axis=lambda e: np.asarray(K.GLOBAL_ROT_INIT, dtype=np.float64))

    class Renderer(torch.nn.Module):                      # stands in for p3d_renderer.Renderer (PyTorch3D 0.2.5)
        def __init__(self, image_size, device):
            super().__init__()
            self.image_size = image_size

        def forward(self, vertices, points, faces, render_texture=False):
            return O.render_silhouettes(state["oracle"], vertices, self.image_size), O.project_points_screen(points, self.image_size)

    stub("p3d_renderer", Renderer=Renderer)
    c = model_io.load_asset()                              # the shipped family-1 asset (== the reference's tables: test_loader_vs_reference)
    m = O.OracleModel.from_constants(c, torch.float32)
    state["oracle"] = m
    data, gt = synthetic.make_sequence(c, N, S, H.oracle_renderer(O.OracleModel.from_constants(c, torch.float64), S), seed=21)
    rgb, sil, joints, vis = data
    p = H.perturbed_params(m, gt, seed=22)

    os.chdir(REF)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "smal_fitter"))
    from smal_fitter import SMALFitter
    import config

    model = SMALFitter("cpu", (rgb.clone(), sil.clone(), joints.clone(), vis.clone()), N, 1, True)
    out = {"sil": np.packbits(sil.numpy().astype(np.uint8)), "joints": joints.numpy(), "vis": vis.numpy(), "S": S, "N": N,
           "init_betas": model.betas.detach().numpy().copy(), "init_log_beta_scales": model.log_beta_scales.detach().numpy().copy(),
           "init_global_rotation": model.global_rotation.detach().numpy().copy()}
    names = ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")
    with torch.no_grad():
        for k in names:
            getattr(model, k).copy_(getattr(p, k))
            out["p_" + k] = getattr(p, k).numpy().copy()
    for stage, weights in enumerate(np.array(config.OPT_WEIGHTS).T[:3]):
        w6, w_temp = [float(x) for x in weights[:6]], float(weights[6])
        br = [1, 2] if stage == 1 else list(range(N))
        for k in names:
            getattr(model, k).grad = None
            getattr(model, k).requires_grad_(True)
        loss, objs = model(br, w6, stage)
        jl, gl, tl = model.get_temporal(w_temp)
        (loss + jl + gl + tl).backward()
        pre = "s%d_" % stage
        out[pre + "weights"] = np.array(w6 + [w_temp])
        out[pre + "batch_range"] = np.array(br)
        out[pre + "loss"] = float(loss)
        for k in ("joint", "sil_reproj", "betas", "pose", "splay"):
            out[pre + "term_" + k] = float(objs[k]) if k in objs else np.nan
        out[pre + "temporal"] = np.array([float(jl), float(gl), float(tl)])
        for k in names:
            out[pre + "grad_" + k] = getattr(model, k).grad.numpy().copy()
    path = os.path.join(HERE, "fitter_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
