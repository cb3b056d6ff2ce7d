"""CPU, world_size 2 over gloo: the frame-sharding arithmetic of the multi-GPU path.
Each rank evaluates the oracle on its contiguous block of frames with the GLOBAL window
normalisers, the shared-shape prior is counted on rank 0 only, gradients are all-reduced
(sum) and the temporal term is added after the reduce -- the result must equal the
unsharded epoch gradient (the identity smalify_b200.smal_fitter.FusedFit relies on)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import helpers as H
    from oracle import smal_oracle as O
    from smalify_b200 import constants as K, model_io, synthetic
    c = model_io.load_asset()
    m = O.OracleModel.from_constants(c, torch.float64)
    S, N = 32, 4
    data, gt = synthetic.make_sequence(c, N, S, H.oracle_renderer(m, S), seed=0)
    rgb, sil, joints, vis = data
    p = H.perturbed_params(m, gt, seed=11)
    w = list(K.STAGE_SCHEDULE[1][:6])
    for t in p.tensors():
        t.requires_grad_(True)
    per = N // world
    lo, hi = rank * per, (rank + 1) * per
    # local part: the frames of this rank, normalised by the global window (N), prior on rank 0 only
    wl = list(w)
    if rank != 0:
        wl[2] = 0.0
    loss, _ = O.fitter_forward(m, p, sil, joints, vis, range(lo, hi), wl, S)
    # fitter_forward normalises by the local count; rescale the per-frame means to the global window
    # (splay is a sum and the prior is per window, so they are handled separately)
    # -> recompute term by term
    for t in p.tensors():
        t.grad = None
    total = torch.zeros((), dtype=torch.float64)
    _, objs = O.fitter_forward(m, p, sil, joints, vis, range(lo, hi), w, S)
    scale = (hi - lo) / N
    for k, v in objs.items():
        if k == "splay":
            total = total + v
        elif k == "betas":
            total = total + (v if rank == 0 else 0.0 * v)
        else:
            total = total + v * scale
    total.backward()
    flat = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1) for t in p.tensors()])
    dist.all_reduce(flat)
    if rank == 0:
        # reference: the unsharded epoch (one window of N frames) without the temporal term
        for t in p.tensors():
            t.grad = None
        full, _ = O.fitter_forward(m, p, sil, joints, vis, range(N), w, S)
        full.backward()
        ref = torch.cat([t.grad.reshape(-1) for t in p.tensors()])
        out.put(float((flat - ref).abs().max() / ref.abs().max()))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_gradient_equals_unsharded():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    err = out.get()
    assert err < 1e-10, err


def _worker_owner_computes(rank, world, port, out):
    """The exchange smalfit_fused_step performs (step_tail_kernel): every rank contributes the gradient rows of ITS
    frames -- including the temporal term's gradient with respect to them, computed from the replicated parameters
    of the neighbouring frames -- and its share of the shared-shape gradient; rows are gathered from their owner,
    shared entries summed in rank order."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import helpers as H
    from oracle import smal_oracle as O
    from smalify_b200 import constants as K, model_io, synthetic
    c = model_io.load_asset()
    m = O.OracleModel.from_constants(c, torch.float64)
    S, N, w_temp = 32, 4, 100.0
    data, gt = synthetic.make_sequence(c, N, S, H.oracle_renderer(m, S), seed=0)
    rgb, sil, joints, vis = data
    p = H.perturbed_params(m, gt, seed=11)
    w = list(K.STAGE_SCHEDULE[1][:6])
    for t in p.tensors():
        t.requires_grad_(True)
    per = N // world
    lo, hi = rank * per, (rank + 1) * per
    _, objs = O.fitter_forward(m, p, sil, joints, vis, range(lo, hi), w, S)
    scale = (hi - lo) / N
    total = torch.zeros((), dtype=torch.float64)
    for k, v in objs.items():
        total = total + (v if k == "splay" else (v if rank == 0 else 0.0 * v) if k == "betas" else v * scale)
    jl, gl, tl = O.temporal_terms(p, w_temp)          # every rank holds all parameters: the whole term, rows of own frames kept
    (total + jl + gl + tl).backward()
    shared = torch.cat([p.betas.grad.reshape(-1), p.log_beta_scales.grad.reshape(-1)])
    dist.all_reduce(shared)                            # summed over ranks
    rows = {}
    for k in ("global_rotation", "joint_rotations", "trans"):
        mine = getattr(p, k).grad[lo:hi].contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)                   # taken from the owner
        rows[k] = torch.cat(parts)
    if rank == 0:
        for t in p.tensors():
            t.grad = None
        full = O.epoch_loss(m, p, sil, joints, vis, N, w, w_temp, S)
        full.backward()
        err = float((shared - torch.cat([p.betas.grad.reshape(-1), p.log_beta_scales.grad.reshape(-1)])).abs().max())
        for k, v in rows.items():
            g = getattr(p, k).grad
            err = max(err, float((v - g).abs().max() / g.abs().max()))
        out.put(err)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_owner_computes_exchange_equals_unsharded_epoch():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_owner_computes, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    err = out.get()
    assert err < 1e-9, err
