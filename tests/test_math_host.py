"""CPU: the closed forms the CUDA kernels use (smalify_b200/csrc/smalfit_math.cuh, compiled for
the host by g++) against the oracle's autograd.  Test infrastructure: the product never loads
this host build."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import smal_oracle as O
from smalify_b200 import constants as K

HERE = os.path.dirname(os.path.abspath(__file__))
fp = ctypes.POINTER(ctypes.c_float)
ip = ctypes.POINTER(ctypes.c_int)
F = lambda a: a.ctypes.data_as(fp)  # noqa: E731
I = lambda a: a.ctypes.data_as(ip)  # noqa: E731


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("chk") / "libcheck_math.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-x", "c++",
                    os.path.join(HERE, "cpu_check", "check_math.cpp"), "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.chk_face_eval.argtypes = [fp, ctypes.c_float, ctypes.c_float, fp, fp]
    return L


def test_rodrigues_forward_backward(lib):
    rng = np.random.default_rng(0)
    for th in [np.zeros(3), rng.normal(size=3) * 0.5, np.array(K.GLOBAL_ROT_INIT), rng.normal(size=3) * 2.0]:
        th32 = th.astype(np.float32)
        R = np.zeros(9, np.float32)
        lib.chk_rodrigues(F(th32), F(R))
        t = torch.tensor(th32, dtype=torch.float64, requires_grad=True)
        Ro = O.rodrigues(t[None])[0]
        Rb = rng.normal(size=(3, 3))
        (Ro * torch.tensor(Rb)).sum().backward()
        thb = np.zeros(3, np.float32)
        Rb32 = Rb.astype(np.float32).reshape(-1).copy()
        lib.chk_rodrigues_bwd(F(th32), F(Rb32), F(thb))
        assert np.abs(R.reshape(3, 3) - Ro.detach().numpy()).max() < 5e-7
        assert np.abs(thb - t.grad.numpy()).max() < 5e-6


def test_chain_forward_backward(lib, constants):
    parents = constants.parents.astype(np.int32)
    sa = constants.tables["scale_axis"].astype(np.int32).reshape(-1)
    rng = np.random.default_rng(1)
    for trial in range(4):
        theta = (rng.normal(size=(35, 3)) * 0.3).astype(np.float32)
        if trial == 0:
            theta[1:] = 0
        J = (rng.normal(size=(35, 3)) * 0.3).astype(np.float32)
        ls = (rng.normal(size=6) * 0.2).astype(np.float32)
        G = np.zeros(35 * 9, np.float32)
        off = np.zeros(35 * 3, np.float32)
        lib.chk_chain(F(theta), F(J), F(ls), I(parents), I(sa), F(G), F(off))
        tt = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        Jt = torch.tensor(J, dtype=torch.float64, requires_grad=True)
        lt = torch.tensor(ls, dtype=torch.float64, requires_grad=True)
        Rs = O.rodrigues(tt).reshape(1, 35, 3, 3)
        _, A = O.global_rigid_transformation(Rs, Jt[None], parents, lt[None])   # the reference's inverse-based loop
        assert np.abs(G.reshape(35, 3, 3) - A[0, :, :3, :3].detach().numpy()).max() < 2e-6
        assert np.abs(off.reshape(35, 3) - A[0, :, :3, 3].detach().numpy()).max() < 2e-6
        Gb = rng.normal(size=(35, 3, 3))
        ob = rng.normal(size=(35, 3))
        ((A[0, :, :3, :3] * torch.tensor(Gb)).sum() + (A[0, :, :3, 3] * torch.tensor(ob)).sum()).backward()
        dth = np.zeros(105, np.float32)
        dJ = np.zeros(105, np.float32)
        dls = np.zeros(6, np.float32)
        Gb32 = Gb.astype(np.float32).reshape(-1).copy()
        ob32 = ob.astype(np.float32).reshape(-1).copy()
        lib.chk_chain_bwd(F(theta), F(J), F(ls), I(parents), I(sa), F(Gb32), F(ob32), F(dth), F(dJ), F(dls))
        for a, b in ((dth.reshape(35, 3), tt.grad.numpy()), (dJ.reshape(35, 3), Jt.grad.numpy()), (dls, lt.grad.numpy())):
            assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max()) + 1e-5


def test_face_eval_decisions_and_gradient(lib):
    rng = np.random.default_rng(2)
    nfrag = 0
    for _ in range(4000):
        ctr = rng.uniform(-0.5, 0.5, size=2)
        tri = np.zeros((3, 3))
        tri[:, :2] = ctr + rng.normal(size=(3, 2)) * 0.03
        tri[:, 2] = 2.5 + rng.normal(size=3) * 0.2
        p = ctr + rng.normal(size=2) * 0.04
        tri32 = tri.astype(np.float32).reshape(-1).copy()
        out = np.zeros(4, np.float32)
        grad = np.zeros(6, np.float32)
        ok = lib.chk_face_eval(F(tri32), np.float32(p[0]), np.float32(p[1]), F(out), F(grad))
        t = torch.tensor(tri32.reshape(3, 3), dtype=torch.float64, requires_grad=True)
        px = torch.tensor(float(np.float32(p[0])), dtype=torch.float64)
        py = torch.tensor(float(np.float32(p[1])), dtype=torch.float64)
        x0, y0, z0, x1, y1, z1, x2, y2, z2 = [t.reshape(-1)[i] for i in range(9)]
        rad = math.sqrt(O.BLUR_RADIUS)
        outb = (px > max(x0, x1, x2) + rad) or (px < min(x0, x1, x2) - rad) or (py > max(y0, y1, y2) + rad) or (py < min(y0, y1, y2) - rad)
        area = O._edge(x2, y2, x0, y0, x1, y1)
        good = (not outb) and not (abs(area) <= O.K_EPS)
        if good:
            den = area + O.K_EPS
            w0 = O._edge(px, py, x1, y1, x2, y2) / den
            w1 = O._edge(px, py, x2, y2, x0, y0) / den
            w2 = O._edge(px, py, x0, y0, x1, y1) / den
            pz = w0 * z0 + w1 * z1 + w2 * z2
            d2 = torch.minimum(torch.minimum(O._seg_dist2(px, py, x0, y0, x1, y1), O._seg_dist2(px, py, x0, y0, x2, y2)),
                               O._seg_dist2(px, py, x1, y1, x2, y2))
            inside = bool((w0 > 0) and (w1 > 0) and (w2 > 0))
            margin = min(abs(float(d2) - O.BLUR_RADIUS), abs(float(pz)))
            if margin < 1e-7 or min(abs(float(w0)), abs(float(w1)), abs(float(w2))) < 1e-4:
                continue                      # borderline decisions may legitimately differ in fp32
            good = bool(pz >= 0) and (inside or bool(d2 < O.BLUR_RADIUS))
        assert good == bool(ok)
        if not good:
            continue
        nfrag += 1
        sd = -d2 if inside else d2
        sd.backward()
        g = t.grad[:, :2].reshape(-1).numpy()
        assert abs(sd.item() - out[1]) < 1e-8
        assert abs(torch.sigmoid(-sd / O.SIGMA).item() - out[2]) < 1e-6
        assert np.abs(g - grad).max() < 1e-6
        wmax = max(abs(float(w0)), abs(float(w1)), abs(float(w2)))
        assert abs(pz.item() - out[0]) < 1e-4 * max(1.0, wmax)
    assert nfrag > 1000


def test_frag_setup_forward_matches_face_eval(lib):
    """The tile rasteriser's fragment test accepts what face_eval (the backward) accepts, with a
    bit-identical depth (the K-nearest threshold is compared across the two) and signed distance."""
    lib.chk_frag_setup_forward.argtypes = [fp, ctypes.c_float, ctypes.c_float, fp]
    rng = np.random.default_rng(11)
    n_both = n_diff = 0
    for _ in range(400):
        ctr = rng.uniform(-0.8, 0.8, size=2)
        tri = np.zeros((3, 3))
        tri[:, :2] = ctr + rng.normal(size=(3, 2)) * rng.choice([0.01, 0.03, 0.1])
        tri[:, 2] = rng.uniform(1.5, 3.5, size=3)
        tri32 = tri.astype(np.float32).reshape(-1).copy()
        out_a, grad, out_b = np.zeros(4, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
        for _ in range(40):
            p = (ctr + rng.normal(size=2) * 0.05).astype(np.float32)
            a = lib.chk_face_eval(F(tri32), p[0], p[1], F(out_a), F(grad))
            b = lib.chk_frag_setup_forward(F(tri32), p[0], p[1], F(out_b))
            if a and b:
                n_both += 1
                assert out_a[0].tobytes() == out_b[0].tobytes()          # depth key, bit for bit
                assert out_a[1].tobytes() == out_b[1].tobytes()          # signed distance too
            else:
                n_diff += int(a != b)
    assert n_both > 2000 and n_diff == 0


def test_face_rect_is_conservative(lib):
    rng = np.random.default_rng(3)
    S = 64
    rect = np.zeros(4, np.int32)
    for _ in range(300):
        ctr = rng.uniform(-1.1, 1.1, size=2)
        tri = np.zeros((3, 3))
        tri[:, :2] = ctr + rng.normal(size=(3, 2)) * 0.05
        tri[:, 2] = 2.5
        tri32 = tri.astype(np.float32).reshape(-1).copy()
        ok = lib.chk_face_rect(F(tri32), S, I(rect))
        out = np.zeros(4, np.float32)
        grad = np.zeros(6, np.float32)
        for r in range(S):
            for c in range(S):
                px, py = np.float32(1 - (2 * c + 1) / S), np.float32(1 - (2 * r + 1) / S)
                if lib.chk_face_eval(F(tri32), px, py, F(out), F(grad)):
                    assert ok and rect[0] <= c <= rect[1] and rect[2] <= r <= rect[3]


def test_camera_backward(lib):
    X = np.array([0.3, -0.2, 0.5], np.float32)
    ndc = np.zeros(3, np.float32)
    lib.chk_camera(F(X), F(ndc))
    t = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    o = O.world_to_ndc(t[None])[0]
    assert np.abs(ndc - o.detach().numpy()).max() < 1e-6
    (o[0] * 0.7 - o[1] * 1.3).backward()
    g3 = np.zeros(3, np.float32)
    g2 = np.array([0.7, -1.3], np.float32)
    lib.chk_camera_bwd(F(ndc), F(g2), F(g3))
    assert np.abs(g3 - t.grad.numpy()).max() < 1e-6
