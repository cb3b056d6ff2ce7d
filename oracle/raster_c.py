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


_LIB64 = os.path.join(_HERE, "_build", "libraster_ref64.so")
_lib64 = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raster_ref.c")
    stale = any(not os.path.exists(l) or os.path.getmtime(l) < os.path.getmtime(src) for l in (_LIB, _LIB64))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True, capture_output=True)
    return _LIB


def _declare(L, suffix, real):
    rp = ctypes.POINTER(real)
    ip = ctypes.POINTER(ctypes.c_int)
    fn = getattr(L, "raster_soft_silhouette" + suffix)
    fn.argtypes = [rp, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, real, real, ctypes.c_int, rp, rp, rp,
                   ctypes.POINTER(ctypes.c_longlong)]
    fn.restype = None
    getattr(L, "raster_num_threads" + suffix).restype = ctypes.c_int
    getattr(L, "raster_set_threads" + suffix).argtypes = [ctypes.c_int]
    getattr(L, "raster_set_threads" + suffix).restype = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _declare(_lib, "", ctypes.c_float)
    return _lib


def lib64():
    global _lib64
    if _lib64 is None:
        build()
        _lib64 = ctypes.CDLL(_LIB64)
        _declare(_lib64, "_f64", ctypes.c_double)
    return _lib64


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double if a.dtype == np.float64 else ctypes.c_float))


def soft_silhouette_np(verts_ndc: np.ndarray, faces: np.ndarray, S: int, mode: int = 1, grad_alpha=None, k=O.K_FACES, dtype=np.float32):
    """verts_ndc (V,3), faces (F,3) i32 -> alpha (S,S) [, grad_verts (V,3)], stats.  dtype float32 (the reference's
    precision) or float64 (libraster_ref64.so)."""
    f64 = np.dtype(dtype) == np.float64
    L = lib64() if f64 else lib()
    fn = L.raster_soft_silhouette_f64 if f64 else L.raster_soft_silhouette
    real = ctypes.c_double if f64 else ctypes.c_float
    v = np.ascontiguousarray(verts_ndc, dtype=dtype)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    alpha = np.empty((S, S), dtype=dtype)
    stats = (ctypes.c_longlong * 4)()
    gv = None
    ga_p = ctypes.POINTER(real)()
    gv_p = ctypes.POINTER(real)()
    if grad_alpha is not None:
        ga = np.ascontiguousarray(grad_alpha, dtype=dtype)
        gv = np.zeros_like(v)
        ga_p, gv_p = _fp(ga), _fp(gv)
    fn(_fp(v), v.shape[0], f.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), f.shape[0], S, k,
       real(O.SIGMA), real(O.BLUR_RADIUS), mode, _fp(alpha), ga_p, gv_p, stats)
    return alpha, gv, dict(n_pair=stats[0], n_frag=stats[1], touched=stats[2], capped=stats[3])


class SoftSilhouetteC(torch.autograd.Function):
    """(B,V,3) NDC verts -> (B,S,S) alpha; backward re-runs the rasteriser with grad_alpha.  float32 or float64
    following the input."""

    @staticmethod
    def forward(ctx, verts_ndc, faces, S, mode):
        dt = np.float64 if verts_ndc.dtype == torch.float64 else np.float32
        v = verts_ndc.detach().numpy().astype(dt, copy=False)
        f = faces.numpy().astype(np.int32)
        out = np.stack([soft_silhouette_np(v[b], f, S, mode, dtype=dt)[0] for b in range(v.shape[0])])
        ctx.save_for_backward(verts_ndc)
        ctx.faces, ctx.S, ctx.mode, ctx.dt = f, S, mode, dt
        return torch.from_numpy(out).to(verts_ndc.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (verts_ndc,) = ctx.saved_tensors
        v = verts_ndc.detach().numpy().astype(ctx.dt, copy=False)
        g = grad_out.numpy().astype(ctx.dt, copy=False)
        gv = np.stack([soft_silhouette_np(v[b], ctx.faces, ctx.S, ctx.mode, grad_alpha=g[b], dtype=ctx.dt)[1] for b in range(v.shape[0])])
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
        lib64().raster_set_threads_f64(n)
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return num_threads()
