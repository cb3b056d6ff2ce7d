"""ctypes wrapper of oracle/raster_ref.c (CPU ORACLE, test / baseline infrastructure).

``SoftSilhouetteC`` is a torch.autograd.Function so the C rasteriser can sit inside
the restated reference CPU path (torch-CPU SMAL + CPU raster) that bench.py times
as ``cpu_baseline`` / ``--impl reference``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

from . import smal_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libraster_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raster_ref.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.raster_soft_silhouette.argtypes = [fp, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_float, ctypes.c_float, ctypes.c_int, fp, fp, fp,
                                                ctypes.POINTER(ctypes.c_longlong)]
        _lib.raster_soft_silhouette.restype = None
        _lib.raster_num_threads.restype = ctypes.c_int
        _lib.raster_set_threads.argtypes = [ctypes.c_int]
        _lib.raster_set_threads.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def soft_silhouette_np(verts_ndc: np.ndarray, faces: np.ndarray, S: int, mode: int = 1, grad_alpha=None, k=O.K_FACES):
    """verts_ndc (V,3) f32, faces (F,3) i32 -> alpha (S,S) [, grad_verts (V,3)], stats."""
    L = lib()
    v = np.ascontiguousarray(verts_ndc, dtype=np.float32)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    alpha = np.empty((S, S), dtype=np.float32)
    stats = (ctypes.c_longlong * 4)()
    gv = None
    ga_p = ctypes.POINTER(ctypes.c_float)()
    gv_p = ctypes.POINTER(ctypes.c_float)()
    if grad_alpha is not None:
        ga = np.ascontiguousarray(grad_alpha, dtype=np.float32)
        gv = np.zeros_like(v)
        ga_p, gv_p = _fp(ga), _fp(gv)
    L.raster_soft_silhouette(_fp(v), v.shape[0], f.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), f.shape[0], S, k,
                             np.float32(O.SIGMA), np.float32(O.BLUR_RADIUS), mode, _fp(alpha), ga_p, gv_p, stats)
    return alpha, gv, dict(n_pair=stats[0], n_frag=stats[1], touched=stats[2], capped=stats[3])


class SoftSilhouetteC(torch.autograd.Function):
    """(B,V,3) NDC verts -> (B,S,S) alpha; backward re-runs the rasteriser with grad_alpha."""

    @staticmethod
    def forward(ctx, verts_ndc, faces, S, mode):
        v = verts_ndc.detach().float().numpy()
        f = faces.numpy().astype(np.int32)
        out = np.stack([soft_silhouette_np(v[b], f, S, mode)[0] for b in range(v.shape[0])])
        ctx.save_for_backward(verts_ndc)
        ctx.faces, ctx.S, ctx.mode = f, S, mode
        return torch.from_numpy(out).to(verts_ndc.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (verts_ndc,) = ctx.saved_tensors
        v = verts_ndc.detach().float().numpy()
        g = grad_out.float().numpy()
        gv = np.stack([soft_silhouette_np(v[b], ctx.faces, ctx.S, ctx.mode, grad_alpha=g[b])[1] for b in range(v.shape[0])])
        return torch.from_numpy(gv).to(verts_ndc.dtype), None, None, None


def num_threads() -> int:
    return int(lib().raster_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every core this process may
    run on (the scheduler affinity mask, not os.cpu_count(): oversubscribing OpenMP is far slower)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if num_threads() < n:
        lib().raster_set_threads(n)
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return num_threads()
