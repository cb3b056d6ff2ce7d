#!/usr/bin/env python
"""bench.py -- fitter iters/sec at WINDOW_SIZE=128, 256x256 silhouettes (BASELINE.json configs[2]).

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
    python bench.py --workload config4 --gpus 8          (BASELINE.json configs[3]: 64 independent 512x512 images per GPU)

One "step" = one epoch of smal_fitter/optimize_to_joints.py:117-137 over the synthetic frames with the
stage-1 weights (every loss term on): forward + analytic backward of all frames, temporal term, [one exchange
of the gradient when frames are sharded], Adam.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from smalify_b200 import constants as K  # noqa: E402
from smalify_b200 import model_io  # noqa: E402

METRIC = "fitter iters/sec at WINDOW_SIZE=128, 256x256 sil"
STAGE = 1      # headline: stage-1 weights, all terms on
QUALITY_FRAMES = (0, 32, 64, 96)          # frames of the 128-frame sequence the quality check fits
QUALITY_ITERS = (30, 40, 20, 10)          # a short 4-stage schedule (config.py:63-72 with fewer epochs)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4"],
                    help="config3: BASELINE.json configs[2], the 128-frame 256x256 sequence sharded over the GPUs (headline); "
                         "config4: configs[3], 64 independent 512x512 images per GPU, one shape each, no collective")
    ap.add_argument("--frames", type=int, default=None, help="frames in total (config3: 128) / per GPU (config4: 64)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-quality", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: bound of the whole CPU run")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: exchange fused into the step-tail kernel over NVLink peer memory, or the unfused NCCL sequence")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def lib_hash() -> str:
    """Identifies the kernels' source (nvcc output is not bit-reproducible: the hash is over csrc/ + include/smalfit.h)."""
    from smalify_b200 import build as B
    return B.source_hash()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def b_alg_bytes(S: int, V: int, F: int) -> float:
    """algorithmic bytes per frame-iteration, 48 V + 24 F + 12 S^2."""
    return 48.0 * V + 24.0 * F + 12.0 * S * S


def workload_of(args):
    if args.workload == "config4":
        return dict(frames_per_gpu=args.frames or 64, S=args.size or 512, per_frame_shapes=True)
    return dict(frames=args.frames or 128, S=args.size or 256, per_frame_shapes=False)


# ---------------------------------------------------------------------------------------------
# CPU side (oracle = checker / baseline only)
# ---------------------------------------------------------------------------------------------
def _oracle_targets(c, n, S, frames=None, n_total=None):
    """Synthetic targets rendered by the oracle's C rasteriser (CPU)."""
    from oracle import cpu_path
    from oracle import smal_oracle as O
    from smalify_b200 import synthetic
    m32 = O.OracleModel.from_constants(c, torch.float32)

    def render(gt):
        k = gt["global_rotation"].shape[0]
        theta = torch.cat([gt["global_rotation"][:, None], gt["joint_rotations"]], 1)
        v, j, _ = O.smal_forward(m32, gt["betas"].expand(k, 20), theta, gt["log_beta_scales"].expand(k, 6))
        v = v + gt["trans"][:, None]
        j = j + gt["trans"][:, None]
        a = cpu_path.c_silhouette_fn(1)(m32, v, S)[:, 0]
        return (a > 0.5).to(torch.uint8), O.project_points_screen(j[:, list(O.CANONICAL)], S).float()
    if frames is None:
        return synthetic.make_sequence(c, n, S, render, seed=0)
    return synthetic.make_subsequence(c, n_total, list(frames), S, render, seed=0)


def run_reference(args):
    """The restated reference CPU path (torch-CPU SMAL + C restatement of the PyTorch3D CPU rasteriser, OpenMP over
    all host cores; PyTorch3D itself is not installable, SMAL half pinned bit-exactly to the reference by
    tests/golden) on the arm's own config: every step is one epoch over ALL frames when the K + W steps fit the
    CPU budget, otherwise over the largest frame sample that does (then `extrapolated` is true and the line says
    so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cpu_path, raster_c
    raster_c.use_all_cores()
    c = model_io.load_asset()
    wl = workload_of(args)
    if args.workload != "config3":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm is defined for the headline workload (config3)"}), flush=True)
        return
    S, N = wl["S"], wl["frames"]
    w = K.STAGE_SCHEDULE[STAGE]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # probe: one epoch on 8 frames (frames are independent in the CPU path: time is linear in the frame count)
    probe_n = min(8, N)
    data_p, _ = _oracle_targets(c, probe_n, S)
    t_probe, _ = cpu_path.time_cpu_epochs(c, data_p, probe_n, w[:6], w[6], w[8], S, 1, mode=1, warmup=1)
    per_frame = t_probe / probe_n
    sample = N
    while sample > 2 and per_frame * sample * (steps + warm) > args.cpu_budget_s:
        sample //= 2
    extrapolated = sample < N
    data, _ = _oracle_targets(c, N, S) if not extrapolated else _oracle_targets(c, sample, S)
    dt, loss = cpu_path.time_cpu_epochs(c, data, sample, w[:6], w[6], w[8], S, steps, mode=1, warmup=warm)
    t_full = dt * (N / sample)
    value = 1.0 / t_full
    cores = raster_c.num_threads()
    sample_txt = (f"all {N} frames per step" if not extrapolated else
                  f"{sample} of {N} frames per step (CPU budget {args.cpu_budget_s:.0f} s), time scaled x{N / sample:.0f} to {N} frames")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "extrapolated": extrapolated,
        "config": {"workload": f"synthetic rs_dog-like sequence, WINDOW_SIZE={N}, {S}x{S} sil, stage-1 weights",
                   "frames": N, "image_size": S, "frames_timed_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_txt}, {steps} steps after {warm} warm-up; torch-CPU SMAL (restated, pinned to the "
                                   f"reference's smal_model by tests/golden) + C/OpenMP restated PyTorch3D rasteriser (culled rows)"},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": loss,
        "note": "reference = restated CPU path (PyTorch3D 0.2.5 not installable offline: parity of the rasteriser half unpinned); "
                "ms_per_step is the measured time of one timed step (the sampled frames)",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(c, N, S, weights, w_temp, lr):
    """cpu_baseline of the default run: the culled OpenMP port on ALL frames (1 warm-up + 1 timed epoch, bounded to
    ~30 s by halving the frame count) and, as BASELINE.md section 4 promises, the faithful mode -- single thread, every
    face against every pixel like RasterizeMeshesNaiveCpu -- on 2 frames, scaled linearly."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cpu_path, raster_c
    raster_c.use_all_cores()
    data_p, _ = _oracle_targets(c, 4, S)
    t_probe, _ = cpu_path.time_cpu_epochs(c, data_p, 4, weights, w_temp, lr, S, 1, mode=1, warmup=1)
    sample = N
    while sample > 4 and (t_probe / 4) * sample * 2 > 30.0:
        sample //= 2
    data, _ = _oracle_targets(c, sample, S)
    dt, _ = cpu_path.time_cpu_epochs(c, data, sample, weights, w_temp, lr, S, 1, mode=1, warmup=1)
    out = {"value": 1.0 / (dt * N / sample), "unit": "iters/s", "cores": raster_c.num_threads(), "kind": "port",
           "sample": (f"{sample} of {N} frames" if sample < N else f"all {N} frames") + ", 1 epoch after 1 warm-up, torch-CPU SMAL + "
                     "C/OpenMP restated PyTorch3D rasteriser (culled rows)" + (f"; time scaled x{N / sample:.0f}" if sample < N else "")}
    try:
        sub = tuple(None if t is None else t[:2] for t in data)
        dt0, _ = cpu_path.time_cpu_epochs(c, sub, 2, weights, w_temp, lr, S, 1, mode=0, warmup=0)
        out["naive_single_thread"] = {"value": 1.0 / (dt0 * N / 2), "unit": "iters/s", "cores": 1,
                                      "sample": f"2 of {N} frames, 1 epoch, every face against every pixel (RasterizeMeshesNaiveCpu-style); "
                                                f"time scaled x{N / 2:.0f}"}
    except Exception as ex:
        out["naive_single_thread"] = {"value": None, "sample": f"failed: {ex!r}"}
    return out


def quality_leg(c, dev):
    """kp-L2 / IoU of a short 4-stage fit on frames QUALITY_FRAMES of the bench sequence at 256x256, GPU (FusedFit,
    CUDA graph) against the oracle running the reference's loop in float32 on the same inputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cpu_path, raster_c
    from oracle import smal_oracle as O
    from smalify_b200 import metrics
    from smalify_b200.optimize_to_joints import fit_sequence
    from smalify_b200.smal_fitter import SMALFitter
    raster_c.use_all_cores()
    S, n = 256, len(QUALITY_FRAMES)
    data, _ = _oracle_targets(c, n, S, frames=QUALITY_FRAMES, n_total=128)
    rgb, sil, joints, vis = data
    t0 = time.perf_counter()
    f = SMALFitter(dev, data, n, 1, True, constants=c)
    fit_sequence(f, K.STAGE_SCHEDULE, n, fused=True, use_graph=True, iters_override=QUALITY_ITERS)
    alpha, kp = f.render()
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    g_kp, g_iou = metrics.keypoint_l2(kp, joints, vis), metrics.silhouette_iou(alpha, sil)
    dropped = int(f.counters()["dropped_bin_entries"])
    t0 = time.perf_counter()
    m = O.OracleModel.from_constants(c, torch.float32)
    p = O.FitParams.initial(m, n, K.GLOBAL_ROT_INIT)
    fn = cpu_path.c_silhouette_fn(1)
    O.fit(m, p, sil, joints, vis, n, K.STAGE_SCHEDULE, S, iters_override=QUALITY_ITERS, silhouette_fn=fn)
    _, _, aux = O.fitter_forward(m, p, sil, joints, vis, range(n), K.STAGE_SCHEDULE[3][:6], S, return_aux=True, silhouette_fn=fn)
    o_kp, o_iou = O.keypoint_l2(aux["proj"], joints, vis), O.silhouette_iou(aux["silhouettes"], sil)
    t_cpu = time.perf_counter() - t0
    return {"kp_l2": g_kp, "iou": g_iou, "oracle_kp_l2": o_kp, "oracle_iou": o_iou,
            "d_kp_l2_vs_oracle": abs(g_kp - o_kp), "d_iou_vs_oracle": abs(g_iou - o_iou),
            "frames": list(QUALITY_FRAMES), "image_size": S, "iters_per_stage": list(QUALITY_ITERS),
            "oracle": "restated reference loop, float32, C rasteriser (CPU)", "gpu_fit_s": t_gpu, "oracle_fit_s": t_cpu,
            "dropped_bin_entries": dropped,
            "note": "identical inputs and schedule; over silhouette stages float32 rounding (sign() of the L1 term, Adam) separates "
                    "any two float32 runs by ~1e-2 px / ~1e-3 IoU (tests/test_gpu_fit_parity.py measures the oracle's own f32-f64 gap)"}


def dropin_leg(fitter_factory, N, weights, w_temp, lr, steps, dev, flush):
    """The reference's epoch, statement for statement (optimize_to_joints.py:113-137), over the drop-in surface:
    SMALFitter.forward + get_temporal + loss.backward() + torch.optim.Adam."""
    model = fitter_factory()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, betas=K.ADAM_BETAS)
    batch_range = list(range(N))

    def epoch():
        acc_loss = 0
        optimizer.zero_grad()
        loss, _ = model(batch_range, weights, STAGE)
        acc_loss = acc_loss + loss.mean()
        joint_loss, global_loss, trans_loss = model.get_temporal(w_temp)
        acc_loss = acc_loss + joint_loss + global_loss + trans_loss
        acc_loss.backward()
        optimizer.step()
        return acc_loss
    for _ in range(3):
        epoch()
    torch.cuda.synchronize()
    ms = []
    for i in range(steps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        epoch()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    for _ in range(steps):
        last = epoch()
    _ = float(last)
    wall = (time.perf_counter() - t0) / steps
    return {"value": 1000.0 / statistics.mean(ms), "unit": "iters/s", "ms_per_step": statistics.mean(ms),
            "wall_ms_per_step_back_to_back": 1000.0 * wall, "steps": steps,
            "api": "SMALFitter.forward + get_temporal + backward + torch.optim.Adam (the reference's loop verbatim)"}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    from smalify_b200 import synthetic
    from smalify_b200.smal_fitter import FusedFit, SMALFitter, _ptr, _stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the fitting path has no CPU implementation)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    wl = workload_of(args)
    S = wl["S"]
    c = model_io.load_asset()
    row = K.STAGE_SCHEDULE[STAGE]
    weights, lr = row[:6], row[8]
    if args.workload == "config4":
        # independent images: every rank fits its own frames_per_gpu frames, nothing is shared, no collective
        per = wl["frames_per_gpu"]
        N = per * world
        lo, hi = rank * per, (rank + 1) * per
        w_temp = 0.0
        data, gt = synthetic.make_subsequence(c, N, list(range(lo, hi)), S, synthetic.gpu_renderer(c, S, dev, per_frame_shapes=True),
                                              seed=0, per_frame_shapes=True)
        torch.cuda.synchronize()
        fitter = SMALFitter(dev, data, 1, 1, True, constants=c, per_frame_shapes=True)
        loop = FusedFit(fitter, 1)
        n_local, frames_rank = per, per
        shard = (0, per)
        scaling = "weak"
    else:
        N = wl["frames"]
        if N % world:
            raise SystemExit(f"--frames {N} must be divisible by the number of ranks {world}")
        w_temp = row[6]
        per = N // world
        lo, hi = rank * per, (rank + 1) * per
        # every rank renders the targets of its own frames only and holds workspace / targets for them only
        data, gt = synthetic.make_subsequence(c, N, list(range(lo, hi)), S, synthetic.gpu_renderer(c, S, dev), seed=0, pad_to=(lo, N))
        torch.cuda.synchronize()
        fitter = SMALFitter(dev, data, N, 1, True, constants=c, frame_shard=(lo, hi))
        loop = FusedFit(fitter, N, process_group=group, collective=args.collective)
        frames_rank = hi - lo
        shard = (lo, hi)
        scaling = "strong"
    mem_after_setup = torch.cuda.memory_allocated(dev)
    free_b, total_b = torch.cuda.mem_get_info(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # kernel launches of one step (counted by the library on an eager step)
    loop.step(weights, w_temp, lr)
    torch.cuda.synchronize()
    c0 = fitter.counters()["launches"]
    loop.step(weights, w_temp, lr)
    torch.cuda.synchronize()
    launches_per_step = fitter.counters()["launches"] - c0

    for _ in range(max(args.warmup, 3)):
        loop.step(weights, w_temp, lr, use_graph=True)
    barrier()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > L2 (126 MB)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                 # untimed L2 flush between timed steps
        starts[i].record()
        loop.step(weights, w_temp, lr, use_graph=True)
        ends[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(ms_steps)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    value = 1000.0 / ms_per_step
    final_loss = float(loop.total_loss())
    fitter.check_faults()

    # replicas: after the timed loop every rank must hold bit-identical parameters and Adam state
    replicas_identical = None
    if world > 1 and args.workload == "config3":
        mine = torch.cat([loop.flat_p, loop.flat_m, loop.flat_v])
        parts = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(parts, mine)
        replicas_identical = all(bool(torch.equal(parts[0], q)) for q in parts[1:])

    # back-to-back steps, one event pair (no flush): informational
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loop.step(weights, w_temp, lr, use_graph=True)
    e1.record()
    barrier()
    b2b = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(b2b, op=torch.distributed.ReduceOp.MAX)
    b2b_ips = 1000.0 * args.steps / float(b2b)

    # ---- e2e: per step, H2D of this rank's targets from pinned host memory + step + D2H loss read.
    # value: the uploads are double-buffered (SMALFitter.stage_targets / swap_targets): step s+1's targets travel on a copy
    # stream under step s's kernels; every step still copies its own inputs inside the timed region (K copies for K steps, the
    # first one not overlapped) and reads its loss back.  serial_value: the same with one set of buffers, copy then step.
    h = fitter._handle
    a0, a1 = shard
    sil_pin, kp_pin = fitter._sil_u8[a0:a1], fitter._joints_f32[a0:a1]
    vis_pin = fitter._vis_u8(0, fitter.num_images).cpu().pin_memory()[a0:a1]
    h2d = (a1 - a0) * (S * S + K.N_KEYPOINTS * 2 * 4 + K.N_KEYPOINTS)

    def e2e_serial(steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            h.check(h.lib.smalfit_set_targets(h.h, a0, a1 - a0, _ptr(sil_pin), _ptr(kp_pin), _ptr(vis_pin), 1, _stream(dev)),
                    "smalfit_set_targets")
            loop.step(weights, w_temp, lr, use_graph=True)
            _ = float(loop.total_loss())          # D2H read of the step's result (syncs)
        barrier()
        return time.perf_counter() - t0

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_pipelined(steps):
        main = torch.cuda.current_stream(dev)
        barrier()
        t0 = time.perf_counter()
        staged = fitter.stage_targets(sil_pin, kp_pin, vis_pin, copy_stream)        # step 0's inputs: nothing to hide under
        for i in range(steps):
            fitter.swap_targets(staged)                                             # this step waits for its own inputs
            if i + 1 < steps:
                free = torch.cuda.Event()
                free.record(main)                                                   # the other set was last read by step i - 1
                copy_stream.wait_event(free)
                staged = fitter.stage_targets(sil_pin, kp_pin, vis_pin, copy_stream)   # step i + 1's inputs, under step i
            loop.step(weights, w_temp, lr, use_graph=True)
            _ = float(loop.total_loss())          # D2H read of the step's result (syncs)
        barrier()
        return time.perf_counter() - t0

    def over_ranks(t):
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        return float(tt)

    e2e_serial_ips = args.steps / over_ranks(e2e_serial(args.steps))
    e2e_pipelined(4)                              # untimed: second set of buffers, one CUDA graph per set
    e2e_ips = args.steps / over_ranks(e2e_pipelined(args.steps))

    # ---- roofline pass: eager steps with per-phase CUDA events (same work, not graph-captured)
    fitter.counters()                        # reset the cumulative pixel counters
    fitter.set_profiling(True)
    prof = []
    n_prof = min(args.steps, 20)
    for _ in range(n_prof):
        flush.fill_(1)
        loop.step(weights, w_temp, lr)
        prof.append(fitter.profile())
    cnt = fitter.counters()
    # one more (untimed) step with the backward counting the pairs it sweeps / uses
    fitter.set_profiling(True, count_pairs=True)
    fitter.work_counts(a0, a1 - a0)          # reset the backward's pair counters
    loop.step(weights, w_temp, lr)
    work = fitter.work_counts(a0, a1 - a0)       # (pixel, face) pairs of this rank's frames in the last pass
    fitter.set_profiling(False)
    phase_ms = {k: statistics.mean(p[k] for p in prof) for k in prof[0]}
    fitter.check_faults()
    peak_fp32 = fitter.fp32_peak() if rank == 0 else None

    def finish():
        # leave without tearing NCCL down: destroying the process group while CUDA graphs that captured
        # its collectives are alive can dead-lock; every measurement is complete at this point
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)

    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        finish()
        return
    peaks, peak_kind = measured_peaks()
    V, F = c.v_template.shape[0], c.faces.shape[0]
    alg_bytes = frames_rank * b_alg_bytes(S, V, F)
    rf_ms = phase_ms["raster_forward"]
    rb_ms = phase_ms["raster_backward"]
    achieved_hbm = alg_bytes / (rf_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture for this build / shape under profiles/"
    tpath = os.path.join(ROOT, "profiles", "raster_forward_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            t = json.load(fh)
        if t.get("frames_per_gpu") == frames_rank and t.get("image_size") == S:
            if t.get("src_sha16") == lib_hash():
                traffic, traffic_note = t["dram_bytes_per_launch"], "ncu --set full capture of this build (profiles/raster_forward_traffic.json)"
            else:
                traffic_note = (f"last capture ({t.get('dram_bytes_per_launch')} B per launch, {t.get('capture', 'earlier build')}) "
                                f"is of other kernel sources: not reported as this build's traffic")
    # The binding roof: FP32 pair tests.  F_alg = 90 flop per bounding-box-passing pair in the forward (70 in the
    # backward); peak = the FMA-chain ceiling measured on this GPU in this run (scalar FFMA; FFMA2 reported beside it).
    pairs = float(work["pairs"])
    fp32_fwd = 90.0 * pairs / (rf_ms * 1e-3) / 1e12
    fp32_bwd = 70.0 * pairs / (rb_ms * 1e-3) / 1e12
    nominal = 148 * 128 * 2 * ((clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) / 1e3) / 1e3
    fp32_peak = peak_fp32["ffma"] if peak_fp32 else nominal
    roofline = {"bound": "fp32", "achieved": fp32_fwd, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_fwd / fp32_peak,
                "peak_source": "FMA-chain microbenchmark in this run (smalfit_fp32_peak, scalar FFMA)" if peak_fp32 else "nominal",
                "peak_ffma2_tflops": peak_fp32["ffma2"] if peak_fp32 else None, "peak_nominal_tflops": nominal,
                "kernel": "raster_tile_forward_kernel", "kernel_ms": rf_ms,
                "algorithmic_flop_per_launch": 90.0 * pairs, "pairs_per_launch": int(pairs), "tile_entries_per_launch": int(work["tile_entries"]),
                "flop_per_pair": {"forward": 90, "backward": 70},
                "traffic": traffic, "traffic_note": traffic_note,
                "backward": {"kernel": "raster_backward_kernel", "kernel_ms": rb_ms, "achieved": fp32_bwd, "frac": fp32_bwd / fp32_peak,
                             "pairs_in_live_pixels_frac": work["live_pairs"] / max(pairs, 1.0),
                             "pairs_contributing_frac": work["used_pairs"] / max(pairs, 1.0)},
                "hbm": {"achieved": achieved_hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_hbm / peaks["hbm_gbs"],
                        "peak_source": peak_kind, "algorithmic_bytes_per_launch": alg_bytes,
                        "note": "north star's HBM fraction (B_alg = 48V + 24F + 12S^2 per frame over the forward kernel's time); the path is "
                                "FP32-ALU / latency bound (SURVEY 8d: ~100 flop/B against a ridge of ~11), so `bound` names the FP32 roof"},
                "phase_ms": phase_ms,
                "step_frac": {"raster_forward": rf_ms / phase_ms["total"], "raster_backward": rb_ms / phase_ms["total"]}}
    if args.workload == "config4":
        metric = "fitter iters/sec, independent images (one shape per frame), 512x512 sil"
        wtxt = (f"BASELINE configs[3]: {N} independent synthetic images ({frames_rank} per GPU), {S}x{S} sil, one shape per image, stage-1 weights "
                f"(kp+sil+pose+shape+splay), Adam step, no collective")
    else:
        metric = METRIC
        wtxt = (f"synthetic rs_dog-like sequence, WINDOW_SIZE={N}, {S}x{S} sil, stage-1 weights "
                f"(kp+sil+pose+shape+splay+temporal), Adam step, frames sharded over {world} GPU(s)")
    line = {
        "metric": metric, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wtxt, "frames": N, "image_size": S, "frames_per_gpu": frames_rank,
                   "parallelism": f"frame-shard x{world}",
                   "collective": ((loop.collective or "none") + (f" (peer unavailable: {loop.peer_error})" if loop.peer_error else "")
                                  + (" fused into the step-tail kernel" if loop.collective == "peer" else "")) if world > 1 else None,
                   "l2": "256 MiB write between timed steps (L2 flush); per-step CUDA events",
                   "cuda_graph": True},
        "e2e": {"value": e2e_ips, "unit": "iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "serial_value": e2e_serial_ips,
                "pipeline": "double-buffered targets (SMALFitter.stage_targets / swap_targets): every step copies its own inputs from pinned "
                            "host memory inside the timed region, step s+1's copy on a copy stream under step s's kernels, the first "
                            "copy not overlapped; serial_value = one buffer set, copy then step"},
        "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
        "clocks": clocks,
        "roofline": roofline,
        "back_to_back_iters_per_s": b2b_ips,
        "wall_s_timed_region": t_wall,
        "final_loss": final_loss,
        "image_iters_per_s": value * N,
        "capped_pixels_per_step": int(cnt["capped_pixels"]) // n_prof, "long_list_pixels_per_step": int(cnt["spilled_pixels"]) // n_prof,
        "dropped_bin_entries": int(cnt["dropped_bin_entries"]),
        "device_memory": {"torch_allocated_bytes": int(mem_after_setup), "device_used_bytes": int(total_b - free_b)},
        "src_sha16": lib_hash(),
    }
    if replicas_identical is not None:
        line["replicas_identical"] = replicas_identical
    assert line["dropped_bin_entries"] == 0, "the (face, tile) pool overflowed: silhouette terms were inexact"
    if world == 1 and args.workload == "config3":
        if not args.no_dropin:
            try:
                full, _ = synthetic.make_subsequence(c, N, list(range(N)), S, synthetic.gpu_renderer(c, S, dev), seed=0)
                line["dropin_api"] = dropin_leg(lambda: SMALFitter(dev, full, N, 1, True, constants=c), N, weights, w_temp, lr,
                                                min(args.steps, 20), dev, flush)
                line["dropin_api"]["fused_over_dropin"] = value / line["dropin_api"]["value"]
            except Exception as ex:
                line["dropin_api"] = {"value": None, "error": repr(ex)}
        if not args.no_quality:
            try:
                line["quality"] = quality_leg(c, dev)
            except Exception as ex:
                line["quality"] = {"error": repr(ex)}
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(c, N, S, weights, w_temp, lr)
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "iters/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex!r}"}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
