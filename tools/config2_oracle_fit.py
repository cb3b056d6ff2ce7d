#!/usr/bin/env python
"""BASELINE.json configs[1] (N = 10 frames, 256x256, the full 4-stage schedule = 1950 Adam steps) fitted by the CPU
oracle -- the reference's loop restated (oracle/smal_oracle.py::fit) over the C restatement of the PyTorch3D CPU
rasteriser -- in float32 (the reference's precision) and in float64, on one set of inputs:

    python tools/config2_oracle_fit.py [--threads T] [--dtype 32|64|both]        (CPU only, ~1 h)

writes  tests/golden/config2_inputs.npz        the inputs (masks, keypoints, visibility), shared with the GPU run
        profiles/r02_config2_oracle_fit.json   final kp-L2 / IoU / loss + per-stage checkpoints of both fits
        profiles/r02_config2_oracle_params.npz final parameters of both fits
tools/run_configs.py (GPU box) fits the same inputs with libsmalfit and reports the three triples side by side."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import cpu_path, raster_c  # noqa: E402
from oracle import smal_oracle as O  # noqa: E402
from smalify_b200 import constants as K, model_io, synthetic  # noqa: E402

N, S = 10, 256
INPUTS = os.path.join(ROOT, "tests", "golden", "config2_inputs.npz")


def make_inputs(c):
    m32 = O.OracleModel.from_constants(c, torch.float32)

    def render(gt):
        k = gt["global_rotation"].shape[0]
        theta = torch.cat([gt["global_rotation"][:, None], gt["joint_rotations"]], 1)
        v, j, _ = O.smal_forward(m32, gt["betas"].expand(k, 20), theta, gt["log_beta_scales"].expand(k, 6))
        v, j = v + gt["trans"][:, None], j + gt["trans"][:, None]
        a = cpu_path.c_silhouette_fn(1)(m32, v, S)[:, 0]
        return (a > 0.5).to(torch.uint8), O.project_points_screen(j[:, list(O.CANONICAL)], S).float()
    (rgb, sil, joints, vis), gt = synthetic.make_sequence(c, N, S, render, seed=0)
    np.savez_compressed(INPUTS, sil=np.packbits(sil.numpy().astype(np.uint8).reshape(N, -1), axis=1), joints=joints.numpy(),
                        visibility=vis.numpy().astype(np.uint8), image_size=S, n_frames=N)


def load_inputs():
    d = np.load(INPUTS)
    n, s = int(d["n_frames"]), int(d["image_size"])
    sil = np.unpackbits(d["sil"], axis=1)[:, :s * s].reshape(n, 1, s, s).astype(np.float32)
    return (torch.zeros(1, 3, 1, 1).expand(n, 3, s, s), torch.from_numpy(sil), torch.from_numpy(d["joints"]).float(),
            torch.from_numpy(d["visibility"].astype(np.float32)))


def fit(c, dtype, log):
    m = O.OracleModel.from_constants(c, dtype)
    rgb, sil, joints, vis = load_inputs()
    p = O.FitParams.initial(m, N, K.GLOBAL_ROT_INIT)
    fn = cpu_path.c_silhouette_fn(1)
    stages = []
    t0 = time.perf_counter()

    def cb(stage, it, loss):
        if it == K.STAGE_SCHEDULE[stage][7] - 1:
            stages.append({"stage": stage, "final_loss": loss, "elapsed_s": time.perf_counter() - t0})
            print(log, stages[-1], flush=True)
    O.fit(m, p, sil.to(dtype), joints.to(dtype), vis, N, K.STAGE_SCHEDULE, S, silhouette_fn=fn, callback=cb)
    _, _, aux = O.fitter_forward(m, p, sil.to(dtype), joints.to(dtype), vis, range(N), K.STAGE_SCHEDULE[3][:6], S, return_aux=True, silhouette_fn=fn)
    res = {"kp_l2": O.keypoint_l2(aux["proj"], joints, vis), "iou": O.silhouette_iou(aux["silhouettes"], sil), "stages": stages,
           "wall_s": time.perf_counter() - t0}
    return res, {k: getattr(p, k).detach().numpy() for k in ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--dtype", default="both")
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    raster_c.lib().raster_set_threads(a.threads)
    raster_c.lib64().raster_set_threads_f64(a.threads)
    c = model_io.load_asset()
    if not os.path.exists(INPUTS):
        make_inputs(c)
    out_json = os.path.join(ROOT, "profiles", "r02_config2_oracle_fit.json")
    out_npz = os.path.join(ROOT, "profiles", "r02_config2_oracle_params.npz")
    out = json.load(open(out_json)) if os.path.exists(out_json) else {}
    params = dict(np.load(out_npz)) if os.path.exists(out_npz) else {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        if a.dtype not in ("both", tag[1:]):
            continue
        res, p = fit(c, dt, tag)
        out[tag] = res
        for k, v in p.items():
            params[f"{tag}_{k}"] = v
        out["config"] = {"frames": N, "image_size": S, "schedule": [list(r) for r in K.STAGE_SCHEDULE], "threads": a.threads,
                         "oracle": "oracle/smal_oracle.py::fit (reference loop restated) + oracle/raster_ref.c (PyTorch3D 0.2.5 CPU rasteriser restated)"}
        with open(out_json, "w") as fh:
            json.dump(out, fh, indent=1)
        np.savez_compressed(out_npz, **params)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
