"""Runs the BASELINE.json configurations that are not the bench line (GPU box) and writes
gpurun_out/configs.json:
  config 2: N=10, 256^2, the full 4-stage schedule (1950 Adam steps), fused + CUDA graph
  config 5: resolution sweep N=32, S in {128..1024}: iters/s with the stage-1 weights, phase times
  parity at 512^2: one frame, loss + gradients against the fp64 oracle
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from smalify_b200 import constants as K, metrics, model_io, synthetic
from smalify_b200.optimize_to_joints import fit_sequence
from smalify_b200.smal_fitter import FusedFit, SMALFitter

out = {}
c = model_io.load_asset()
dev = torch.device("cuda", 0)

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

# ---- config 5: resolution sweep --------------------------------------------------------------
sweep = []
for S in (128, 256, 384, 512, 768, 1024):
    N = 32
    data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=0)
    f = SMALFitter(dev, data, N, 1, True, constants=c)
    loop = FusedFit(f, N)
    ms = timed_steps(loop, K.STAGE_SCHEDULE[1], 20)
    f.set_profiling(True)
    loop.step(K.STAGE_SCHEDULE[1][:6], K.STAGE_SCHEDULE[1][6], K.STAGE_SCHEDULE[1][8])
    prof = f.profile(); f.set_profiling(False)
    cnt = f.counters()
    sweep.append(dict(S=S, N=N, ms_per_iter=ms, iters_per_s=1000.0 / ms, phase_ms=prof, counters=cnt))
    print("sweep", sweep[-1], flush=True)
    del loop, f
out["config5_resolution_sweep"] = sweep

# ---- config 2: full 4-stage fit, N=10 ---------------------------------------------------------
N, S = 10, 256
data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=0)
f = SMALFitter(dev, data, N, 1, True, constants=c)
torch.cuda.synchronize(); t0 = time.perf_counter()
finals = fit_sequence(f, K.STAGE_SCHEDULE, N, fused=True, use_graph=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
alpha, kp = f.render()
res = dict(N=N, S=S, iterations=sum(r[7] for r in K.STAGE_SCHEDULE), wall_s=dt, iters_per_s=sum(r[7] for r in K.STAGE_SCHEDULE) / dt,
           final_stage_losses=finals, kp_l2_px=metrics.keypoint_l2(kp, data[2], data[3]),
           iou=metrics.silhouette_iou(alpha, data[1]), counters=f.counters())
gt_kp = None
print("config2", res, flush=True)
out["config2_full_fit"] = res

# ---- parity at 512^2 (one frame) ---------------------------------------------------------------
try:
    import helpers as H
    from oracle import smal_oracle as O
    m = O.OracleModel.from_constants(c, torch.float64)
    S, N = 512, 1
    data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=1)
    p = H.perturbed_params(m, gt, seed=2, scale=0.5)
    w = K.STAGE_SCHEDULE[1][:6]
    t0 = time.perf_counter()
    lo, objs_o, go = H.oracle_loss_and_grads(m, p, data, range(N), w, S)
    f = SMALFitter(dev, data, N, 1, True, constants=c)
    H.load_params_into(f, p)
    loss, objs = f(list(range(N)), w, 1)
    loss.backward()
    par = dict(S=S, loss_gpu=float(loss), loss_oracle=lo, oracle_s=time.perf_counter() - t0,
               grad_rel_err={k: H.rel_err(getattr(f, k).grad, go[k]) for k in go}, counters=f.counters())
    print("parity512", par, flush=True)
    out["parity_512"] = par
except Exception as ex:
    out["parity_512"] = {"error": repr(ex)}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as fh:
    json.dump(out, fh, indent=1, default=float)
print("done")
