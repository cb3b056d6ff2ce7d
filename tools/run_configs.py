"""Runs the BASELINE.json configurations that are not the bench line (GPU box) and writes gpurun_out/configs.json:
  config 2: N=10, 256^2, the full 4-stage schedule (1950 Adam steps) on the inputs of tests/golden/config2_inputs.npz,
            fused + CUDA graph AND through the drop-in surface; final kp-L2 / IoU next to the CPU oracle's float32 and
            float64 fits of the same inputs (profiles/r02_config2_oracle_fit.json, made by tools/config2_oracle_fit.py)
  config 5: resolution sweep N=32, S in {128..1024}: iters/s with the stage-1 weights, phase times
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import torch
from smalify_b200 import constants as K, metrics, model_io, synthetic
from smalify_b200.optimize_to_joints import fit_sequence
from smalify_b200.smal_fitter import FusedFit, SMALFitter

out = {}
c = model_io.load_asset()
dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["config2", "config5"]


def timed_steps(loop, row, steps, graph=True):
    w, wt, lr = row[:6], row[6], row[8]
    for _ in range(4):
        loop.step(w, wt, lr, use_graph=graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loop.step(w, wt, lr, use_graph=graph)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if "config5" in which:
    sweep = []
    for S in (128, 256, 384, 512, 768, 1024):
        N = 32
        data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=0)
        f = SMALFitter(dev, data, N, 1, True, constants=c)
        loop = FusedFit(f, N)
        ms = timed_steps(loop, K.STAGE_SCHEDULE[1], 20)
        f.counters()
        f.set_profiling(True)
        loop.step(K.STAGE_SCHEDULE[1][:6], K.STAGE_SCHEDULE[1][6], K.STAGE_SCHEDULE[1][8])
        prof = f.profile(); work = f.work_counts(); f.set_profiling(False)
        cnt = f.counters()
        f.check_faults()
        sweep.append(dict(S=S, N=N, ms_per_iter=ms, iters_per_s=1000.0 / ms, phase_ms=prof, counters=cnt, work=work,
                          fp32_forward_tflops=90.0 * work["pairs"] / (prof["raster_forward"] * 1e-3) / 1e12,
                          device_memory_bytes=int(torch.cuda.memory_allocated(dev))))
        print("sweep", sweep[-1], flush=True)
        del loop, f
    out["config5_resolution_sweep"] = sweep

if "config2" in which:
    import config2_oracle_fit as C2
    data = C2.load_inputs()
    N, S = data[1].shape[0], data[1].shape[-1]
    res = {}
    for mode in ("fused_graph", "dropin"):
        f = SMALFitter(dev, data, N, 1, True, constants=c)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        finals = fit_sequence(f, K.STAGE_SCHEDULE, N, fused=(mode == "fused_graph"), use_graph=(mode == "fused_graph"))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        alpha, kp = f.render()
        iters = sum(r[7] for r in K.STAGE_SCHEDULE)
        res[mode] = dict(iterations=iters, wall_s=dt, iters_per_s=iters / dt, final_stage_losses=finals,
                         kp_l2=metrics.keypoint_l2(kp, data[2], data[3]), iou=metrics.silhouette_iou(alpha, data[1]), counters=f.counters())
        f.check_faults()
        print("config2", mode, res[mode], flush=True)
    cmp_ = dict(N=N, S=S, gpu=res)
    opath = os.path.join(ROOT, "profiles", "r02_config2_oracle_fit.json")
    if os.path.exists(opath):
        o = json.load(open(opath))
        for tag in ("f32", "f64"):
            if tag in o:
                cmp_["oracle_" + tag] = {k: o[tag][k] for k in ("kp_l2", "iou", "wall_s")}
                cmp_["oracle_" + tag]["final_stage_losses"] = [s["final_loss"] for s in o[tag]["stages"]]
        ref = cmp_.get("oracle_f64") or cmp_.get("oracle_f32")
        if ref:
            for mode in res:
                cmp_[f"d_{mode}_vs_oracle"] = {"kp_l2": abs(res[mode]["kp_l2"] - ref["kp_l2"]), "iou": abs(res[mode]["iou"] - ref["iou"])}
            if "oracle_f64" in cmp_ and "oracle_f32" in cmp_:
                cmp_["d_oracle_f32_vs_f64"] = {"kp_l2": abs(cmp_["oracle_f32"]["kp_l2"] - cmp_["oracle_f64"]["kp_l2"]),
                                               "iou": abs(cmp_["oracle_f32"]["iou"] - cmp_["oracle_f64"]["iou"])}
    print("config2 comparison", json.dumps(cmp_, default=float), flush=True)
    out["config2_full_fit"] = cmp_

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as fh:
    json.dump(out, fh, indent=1, default=float)
print("done")
