#!/usr/bin/env python
"""Multi-GPU check of libsmalfit's one-shot peer-memory all-reduce (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_check.py

1. random vectors: smalfit_peer_allreduce == the rank-ordered sum, bit for bit, and == NCCL all_reduce to rounding,
   eager and replayed from a CUDA graph, identical on every rank;
2. latency of both collectives on the fitter's payload (CUDA events, max over ranks);
3. a few FusedFit epochs -- with the exchange fused into the step-tail kernel (smalfit_fused_step, "peer"), with the
   unfused NCCL sequence, and unsharded on one GPU -- end at the same loss and parameters; the replicas of the fused
   run are bit-identical on every rank.
Prints one JSON line on rank 0; exit code 1 on any mismatch.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from smalify_b200 import constants as K, model_io, synthetic  # noqa: E402
from smalify_b200.smal_fitter import FusedFit, SMALFitter, _ptr, _stream  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    c = model_io.load_asset()
    N, S = 8 * world, 64
    data, gt = synthetic.make_sequence(c, N, S, synthetic.gpu_renderer(c, S, dev), seed=0)
    per = N // world
    out = {"world": world}
    ok = True

    def make(collective):
        # the handle only holds this rank's frames (workspace and targets O(N / ranks))
        f = SMALFitter(dev, data, N, 1, True, constants=c, frame_shard=(rank * per, (rank + 1) * per))
        return f, FusedFit(f, N, process_group=dist.group.WORLD, collective=collective)

    f_peer, loop_peer = make("peer")
    assert loop_peer.collective == "peer"
    h = f_peer._handle
    n = loop_peer.flat_g.numel()
    # 1. random vectors
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(6):
        x = torch.randn(n, device=dev, generator=gen)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        want = parts[0].clone()
        for r in range(1, world):
            want += parts[r]                      # rank order, as the kernel sums
        nccl = x.clone()
        dist.all_reduce(nccl)
        y = x.clone()
        if it < 3:
            h.check(h.lib.smalfit_peer_allreduce(h.h, _ptr(y), n, _stream(dev)), "smalfit_peer_allreduce")
        else:                                     # from a CUDA graph
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            ybuf = x.clone()
            with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
                h.check(h.lib.smalfit_peer_allreduce(h.h, _ptr(ybuf), n, _stream(dev)), "smalfit_peer_allreduce")
            torch.cuda.current_stream(dev).wait_stream(s)
            g.replay()
            y = ybuf
        torch.cuda.synchronize()
        ok &= bool(torch.equal(y, want))
        ok &= bool(torch.allclose(y, nccl, rtol=1e-5, atol=1e-5))
    ok &= not loop_peer.peer_timed_out()
    # 2. latency
    x = torch.randn(n, device=dev)
    lat = {}
    for name, fn in (("peer", lambda: h.lib.smalfit_peer_allreduce(h.h, _ptr(x), n, _stream(dev))), ("nccl", lambda: dist.all_reduce(x))):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lat[name] = float(t) * 1e3
        x.normal_()
    out["latency_us"] = lat
    out["payload_bytes"] = n * 4
    # 3. a few epochs with either collective
    f_nccl, loop_nccl = make("nccl")
    row = K.STAGE_SCHEDULE[1]
    finals = {}
    for name, loop in (("peer", loop_peer), ("nccl", loop_nccl)):
        loop.reset_optimizer()
        for _ in range(12):
            loop.step(row[:6], row[6], row[8], use_graph=True)
        torch.cuda.synchronize()
        finals[name] = float(loop.total_loss())
    # the same epochs unsharded on this GPU alone
    f_one = SMALFitter(dev, data, N, 1, True, constants=c)
    loop_one = FusedFit(f_one, N)
    loop_one.reset_optimizer()
    for _ in range(12):
        loop_one.step(row[:6], row[6], row[8], use_graph=True)
    torch.cuda.synchronize()
    finals["one_gpu"] = float(loop_one.total_loss())
    out["final_loss"] = finals
    ok &= abs(finals["peer"] - finals["nccl"]) <= 1e-4 * abs(finals["nccl"])
    ok &= abs(finals["peer"] - finals["one_gpu"]) <= 1e-4 * abs(finals["one_gpu"])
    dp = float((loop_peer.flat_p - loop_one.flat_p).abs().max())
    dn = float((loop_peer.flat_p - loop_nccl.flat_p).abs().max())
    out["max_param_diff"] = {"peer_vs_one_gpu": dp, "peer_vs_nccl": dn}
    ok &= dp < 1e-4 and dn < 1e-4
    # replicas of the fused run: bit-identical parameters and Adam state on every rank
    mine = torch.cat([loop_peer.flat_p, loop_peer.flat_m, loop_peer.flat_v])
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out["replicas_identical"] = all(bool(torch.equal(parts[0], q)) for q in parts[1:])
    ok &= out["replicas_identical"]
    ok &= not loop_peer.peer_timed_out()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    os._exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
