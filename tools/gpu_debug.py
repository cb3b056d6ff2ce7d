"""First-light diagnostics on the GPU box: prints every oracle-vs-kernel difference
(no asserts), writes gpurun_out/debug.log."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import smal_oracle as O
from smalify_b200 import model_io, synthetic, constants as K
from smalify_b200.smal_fitter import SMALFitter
import helpers as H

def main():
    c = model_io.load_asset()
    m = O.OracleModel.from_constants(c, torch.float64)
    S, N = 64, 3
    data, gt = synthetic.make_sequence(c, N, S, H.oracle_renderer(m, S), seed=0)
    f = SMALFitter("cuda", data, N, 1, True, constants=c)
    print("created fitter", flush=True)
    states = {"init": O.FitParams.initial(m, N, K.GLOBAL_ROT_INIT), "mid": H.perturbed_params(m, gt, 5)}
    for name, p in states.items():
        H.load_params_into(f, p)
        v = f.vertices().cpu().double()
        theta = torch.cat([p.global_rotation[:, None], p.joint_rotations], 1)
        vo, jo, _ = O.smal_forward(m, p.betas.expand(N, 20), theta, p.log_beta_scales.expand(N, 6))
        vo = vo + p.trans[:, None]; jo = jo + p.trans[:, None]
        print(name, "verts maxerr", float((v - vo).abs().max()), flush=True)
        alpha, kp = f.render()
        torch.cuda.synchronize()
        kpo = O.project_points_screen(jo[:, list(O.CANONICAL)], S)
        print(name, "kp maxerr px", float((kp.cpu().double() - kpo).abs().max()))
        ao = O.render_silhouettes(m, vo, S)
        err = (alpha.cpu().double() - ao).abs()
        print(name, "alpha err max", float(err.max()), "mean", float(err.mean()), "frac>2e-5", float((err > 2e-5).double().mean()),
              "sum", float(alpha.sum()), float(ao.sum()), flush=True)
        print(name, "counters", f.counters())
        for label, w in (("stage0", K.STAGE_SCHEDULE[0][:6]), ("stage1", K.STAGE_SCHEDULE[1][:6])):
            lo, objs_o, go = H.oracle_loss_and_grads(m, p, data, range(N), w, S)
            H.load_params_into(f, p)
            for t in f.parameters():
                t.grad = None; t.requires_grad_(True)
            loss, objs = f(list(range(N)), w, 1)
            loss.backward()
            torch.cuda.synchronize()
            print(name, label, "loss", float(loss), lo, {k: (float(objs[k]), objs_o[k]) for k in objs_o})
            for k in go:
                g = getattr(f, k).grad
                print("   grad", k, "rel err", H.rel_err(g, go[k]) if g is not None else None, "scale", float(go[k].abs().max()))
    # timing at benchmark size
    try:
        S2, N2 = 256, 128
        render = synthetic.gpu_renderer(c, S2)
        data2, gt2 = synthetic.make_sequence(c, N2, S2, render, seed=0)
        f2 = SMALFitter("cuda", data2, N2, 1, True, constants=c)
        from smalify_b200.smal_fitter import FusedFit
        ff = FusedFit(f2, N2)
        w = K.STAGE_SCHEDULE[1]
        for use_graph in (False, True):
            for _ in range(3):
                ff.step(w[:6], w[6], w[8], use_graph=use_graph)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(10):
                ff.step(w[:6], w[6], w[8], use_graph=use_graph)
            torch.cuda.synchronize()
            print("bench N=128 S=256 graph", use_graph, "ms/iter", (time.time() - t0) * 100, "loss", float(ff.total_loss()), flush=True)
        print("counters", f2.counters())
    except Exception:
        traceback.print_exc()

if __name__ == "__main__":
    main()
