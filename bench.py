#!/usr/bin/env python
"""bench.py -- fitter iters/sec at WINDOW_SIZE=128, 256x256 silhouettes (BASELINE.json configs[2]).

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1

One "step" = one epoch of smal_fitter/optimize_to_joints.py:117-137 over the 128 synthetic
frames with the stage-1 weights (every loss term on): forward + analytic backward of all
frames, temporal term, [one all-reduce of the flat gradient when frames are sharded], Adam.
Prints ONE JSON line (rank 0).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=8)
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: libsmalfit's one-shot all-reduce over NVLink peer memory, or NCCL")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


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


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The restated reference CPU path (torch-CPU SMAL + C restatement of the PyTorch3D CPU
    rasteriser, OpenMP over all host cores) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cpu_path, raster_c
    from oracle import smal_oracle as O
    import helpers as H
    raster_c.use_all_cores()
    c = model_io.load_asset()
    S, N = args.size, args.frames
    sample = max(2, min(args.cpu_sample_frames, N))
    m64 = O.OracleModel.from_constants(c, torch.float32)
    from smalify_b200 import synthetic
    # targets for the sample frames: rendered by the oracle itself (C rasteriser)
    def render(gt):
        n = gt["global_rotation"].shape[0]
        theta = torch.cat([gt["global_rotation"][:, None], gt["joint_rotations"]], 1)
        v, j, _ = O.smal_forward(m64, gt["betas"].expand(n, 20), theta, gt["log_beta_scales"].expand(n, 6))
        v = v + gt["trans"][:, None]; j = j + gt["trans"][:, None]
        a = cpu_path.c_silhouette_fn(1)(m64, v, S)[:, 0]
        return (a > 0.5).to(torch.uint8), O.project_points_screen(j[:, list(O.CANONICAL)], S).float()
    data, _ = synthetic.make_sequence(c, sample, S, render, seed=0)
    w = K.STAGE_SCHEDULE[STAGE]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bound the run: ~1 s per sampled frame-epoch at 256^2 -> cap the number of epochs
    t_probe, _ = cpu_path.time_cpu_epochs(c, data, sample, w[:6], w[6], w[8], S, 1, mode=1, warmup=0)
    budget_s = 150.0
    steps = max(1, min(steps, int(budget_s / max(t_probe, 1e-3))))
    warm = min(warm, 2)
    dt, loss = cpu_path.time_cpu_epochs(c, data, sample, w[:6], w[6], w[8], S, steps, mode=1, warmup=warm)
    t_full = dt * (N / sample)               # frames are independent: linear in the frame count
    value = 1.0 / t_full
    cores = raster_c.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 * t_full, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic rs_dog-like sequence, WINDOW_SIZE={N}, {S}x{S} sil, stage-1 weights",
                   "frames": N, "image_size": S},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of {N} frames per step, {steps} steps, torch-CPU SMAL + C/OpenMP restated "
                                   f"PyTorch3D rasteriser (culled rows), time scaled x{N / sample:.0f} to {N} frames"},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = restated CPU path (PyTorch3D not installable: parity of the rasteriser half unpinned)",
    }
    print(json.dumps(line), flush=True)


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
    N, S = args.frames, args.size
    if N % world:
        raise SystemExit(f"--frames {N} must be divisible by the number of ranks {world}")
    c = model_io.load_asset()
    data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=0)
    torch.cuda.synchronize()
    per = N // world
    lo, hi = rank * per, (rank + 1) * per
    fitter = SMALFitter(dev, data, N, 1, True, constants=c)
    loop = FusedFit(fitter, N, frame_shard=(lo, hi), process_group=group, collective=args.collective if world > 1 else "nccl")
    row = K.STAGE_SCHEDULE[STAGE]
    weights, w_temp, lr = row[:6], row[6], row[8]

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

    # restart the fit so the timed steps run on the early (largest silhouette error) part
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

    # ---- e2e: per step, H2D of this rank's targets from pinned host memory + step + D2H loss read
    h = fitter._handle
    sil_pin, kp_pin = fitter._sil_u8, fitter._joints_f32
    vis_pin = fitter._vis_u8(0, N).cpu().pin_memory()
    h2d = (hi - lo) * (S * S + K.N_KEYPOINTS * 2 * 4 + K.N_KEYPOINTS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.check(h.lib.smalfit_set_targets(h.h, lo, hi - lo, _ptr(sil_pin[lo:hi]), _ptr(kp_pin[lo:hi]), _ptr(vis_pin[lo:hi]),
                                          1, _stream(dev)), "smalfit_set_targets")
        loop.step(weights, w_temp, lr, use_graph=True)
        _ = float(loop.total_loss())          # D2H read of the step's result (syncs)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(e2e_t, op=torch.distributed.ReduceOp.MAX)
    e2e_ips = args.steps / float(e2e_t)

    # ---- roofline pass: eager steps with per-phase CUDA events (same work, not graph-captured)
    fitter.set_profiling(True)
    prof = []
    for _ in range(min(args.steps, 20)):
        flush.fill_(1)
        loop.step(weights, w_temp, lr)
        prof.append(fitter.profile())
    fitter.set_profiling(False)
    phase_ms = {k: statistics.mean(p[k] for p in prof) for k in prof[0]}
    cnt = fitter.counters()
    work = fitter.work_counts(lo, hi - lo)       # (pixel, face) pairs of this rank's frames in the last pass

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
    frames_rank = hi - lo
    alg_bytes = frames_rank * b_alg_bytes(S, V, F)
    rf_ms = phase_ms["raster_forward"]
    achieved = alg_bytes / (rf_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "raster_forward_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            t = json.load(fh)
        if t.get("frames_per_gpu") == frames_rank and t.get("image_size") == S:
            traffic = t["dram_bytes_per_launch"]      # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full
    # the physically binding roof: FP32 pair tests.  F_alg = 90 flop per bounding-box-passing pair in the
    # forward (+ 70 per pair in the backward); peak = SMs x 128 lanes x 2 x the SM clock seen during the run.
    sm_clock_ghz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) / 1e3
    fp32_peak = 148 * 128 * 2 * sm_clock_ghz / 1e3          # TFLOP/s
    fp32_fwd = 90.0 * work["pairs"] / (rf_ms * 1e-3) / 1e12
    fp32_bwd = 70.0 * work["pairs"] / (phase_ms["raster_backward"] * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                "kernel": "raster_tile_forward_kernel", "kernel_ms": rf_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "phase_ms": phase_ms,
                "fp32": {"pairs_per_launch": int(work["pairs"]), "tile_entries_per_launch": int(work["tile_entries"]),
                         "flop_per_pair": {"forward": 90, "backward": 70}, "peak_tflops": fp32_peak,
                         "forward_tflops": fp32_fwd, "forward_frac": fp32_fwd / fp32_peak,
                         "backward_tflops": fp32_bwd, "backward_frac": fp32_bwd / fp32_peak},
                "note": "path is FP32-ALU / latency bound (SURVEY 8d): the HBM fraction is reported as the contract asks, "
                        "the fp32 block is the binding roof"}
    line = {
        "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic rs_dog-like sequence, WINDOW_SIZE={N}, {S}x{S} sil, stage-1 weights "
                               f"(kp+sil+pose+shape+splay+temporal), Adam step, frames sharded over {world} GPU(s)",
                   "frames": N, "image_size": S, "frames_per_gpu": frames_rank, "parallelism": f"frame-shard x{world}",
                   "collective": (loop.collective + (" TIMED OUT" if loop.collective == "peer" and loop.peer_timed_out() else "")
                                  + (f" (peer unavailable: {loop.peer_error})" if getattr(loop, "peer_error", None) else "")) if world > 1 else None,
                   "l2": "256 MiB write between timed steps (L2 flush); per-step CUDA events",
                   "cuda_graph": True},
        "e2e": {"value": e2e_ips, "unit": "iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "roofline": roofline,
        "back_to_back_iters_per_s": b2b_ips,
        "wall_s_timed_region": t_wall,
        "final_loss": final_loss,
        "capped_pixels_last_pass": int(cnt["capped_pixels"]), "spilled_pixels_last_pass": int(cnt["spilled_pixels"]),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle import cpu_path, raster_c
            raster_c.use_all_cores()
            sample = max(2, min(args.cpu_sample_frames, N))
            sub = tuple(None if t is None else t[:sample] for t in data)
            dt, _ = cpu_path.time_cpu_epochs(c, sub, sample, weights, w_temp, lr, S, 2, mode=1, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / (dt * N / sample), "unit": "iters/s", "cores": raster_c.num_threads(),
                                    "kind": "port",
                                    "sample": f"first {sample} of {N} frames, 2 epochs after 1 warm-up, torch-CPU SMAL + C/OpenMP "
                                              f"restated PyTorch3D rasteriser; time scaled x{N / sample:.0f}"}
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
