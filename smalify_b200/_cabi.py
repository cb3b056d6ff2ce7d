"""ctypes binding of libsmalfit.so (include/smalfit.h).

This is the stub a SMALify maintainer would add next to ``smal_fitter.py``: raw
device pointers of torch tensors and the current CUDA stream are handed through
the C-ABI.  There is no fallback: if the library is missing or fails to load,
importing the hot path raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMALFIT_LIB points tools/ab_bench.py at an experiment build of the same library (never a fallback:
# a missing file raises below)
LIB_PATH = os.environ.get("SMALFIT_LIB") or os.path.join(_HERE, "libsmalfit.so")

ABI_VERSION = 4
L_JOINT, L_SIL, L_BETAS, L_POSE, L_LIMIT, L_SPLAY, L_TEMPORAL, L_TOTAL = range(8)
N_TERMS_FUSED = 12          # smalfit_fused_step: + (joint, global, trans) temporal values at [8..10]
STATUS_POOL_OVERFLOW, STATUS_PEER_TIMEOUT = 1, 2

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_verts", C.c_int32), ("n_faces", C.c_int32),
        ("v_template", _f32p), ("shapedirs", _f32p), ("faces", _i32p), ("parents", _i32p), ("scale_axis", _i32p),
        ("skin_joint", _i32p), ("skin_weight", _f32p),
        ("skinT_ptr", _i32p), ("skinT_vert", _i32p), ("skinT_weight", _f32p),
        ("jreg_ptr", _i32p), ("jreg_vert", _i32p), ("jreg_weight", _f32p),
        ("jregT_ptr", _i32p), ("jregT_joint", _i32p), ("jregT_weight", _f32p),
        ("mj_ptr", _i32p), ("mj_vert", _i32p), ("mj_weight", _f32p),
        ("mjT_ptr", _i32p), ("mjT_joint", _i32p), ("mjT_weight", _f32p),
        ("v2f_ptr", _i32p), ("v2f_fc", _i32p), ("keypoint_joint", _i32p),
        ("pose_mean", _f32p), ("pose_prec", _f32p), ("pose_use", _f32p),
        ("shape_dim", C.c_int32), ("shape_mean", _f32p), ("shape_prec", _f32p),
    ]


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("frame_base", C.c_int32), ("frame_capacity", C.c_int32),
                ("pool_entries_per_frame", C.c_int32)]


class Tensors(C.Structure):
    _fields_ = [("betas", C.c_void_p), ("log_beta_scales", C.c_void_p), ("global_rotation", C.c_void_p),
                ("joint_rotations", C.c_void_p), ("trans", C.c_void_p)]


class SmalfitError(RuntimeError):
    pass


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libsmalfit.so, check the ABI version and declare every prototype."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SmalfitError(
            f"{path} not found: build it with `python -m smalify_b200.build` (needs nvcc). "
            "The SMAL fitting hot path has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    lib.smalfit_abi_version.restype = C.c_int
    if lib.smalfit_abi_version() != ABI_VERSION:
        raise SmalfitError(f"libsmalfit ABI {lib.smalfit_abi_version()} != binding ABI {ABI_VERSION}")
    vp = C.c_void_p
    TP = C.POINTER(Tensors)
    protos = {
        "smalfit_create": ([C.POINTER(ModelDesc), C.c_int, C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "smalfit_create_ex": ([C.POINTER(ModelDesc), C.c_int, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)], C.c_int),
        "smalfit_destroy": ([vp], None),
        "smalfit_status": ([vp, C.POINTER(C.c_int)], C.c_int),
        "smalfit_fused_step": ([vp, TP, TP, TP, TP, C.c_int, C.c_int, C.c_int, _f32p, C.c_float, C.c_int, _i32p,
                                C.c_float, C.c_float, C.c_float, C.c_float, vp, vp], C.c_int),
        "smalfit_fp32_peak": ([vp, _f32p, vp], C.c_int),
        "smalfit_last_error": ([vp], C.c_char_p),
        "smalfit_set_targets": ([vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp], C.c_int),
        "smalfit_stage_targets": ([vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp], C.c_int),
        "smalfit_swap_targets": ([vp, C.POINTER(C.c_int)], C.c_int),
        "smalfit_set_visibility": ([vp, C.c_int, C.c_int, vp, C.c_int, vp], C.c_int),
        "smalfit_set_masks": ([vp, _f32p, _f32p], C.c_int),
        "smalfit_set_windows": ([vp, _i32p, C.c_int], C.c_int),
        "smalfit_set_joint_limits": ([vp, _f32p, _f32p], C.c_int),
        "smalfit_set_focal": ([vp, vp, vp], C.c_int),
        "smalfit_set_per_frame_shapes": ([vp, C.c_int], C.c_int),
        "smalfit_loss_grad": ([vp, TP, C.c_int, C.c_int, _f32p, C.c_int, TP, vp, vp], C.c_int),
        "smalfit_temporal": ([vp, TP, C.c_int, C.c_float, TP, vp, vp], C.c_int),
        "smalfit_adam_step": ([vp, TP, TP, TP, TP, C.c_int, _i32p, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_int, vp], C.c_int),
        "smalfit_adam_reset": ([vp, vp], C.c_int),
        "smalfit_render": ([vp, TP, C.c_int, C.c_int, vp, vp, vp], C.c_int),
        "smalfit_vertices": ([vp, TP, C.c_int, C.c_int, vp, vp], C.c_int),
        "smalfit_set_profiling": ([vp, C.c_int], C.c_int),
        "smalfit_get_profile": ([vp, _f32p], C.c_int),
        "smalfit_render_color": ([vp, vp, C.c_int, _f32p, vp, vp], C.c_int),
        "smalfit_peer_init": ([vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte)], C.c_int),
        "smalfit_peer_connect": ([vp, C.POINTER(C.c_ubyte)], C.c_int),
        "smalfit_peer_allreduce": ([vp, vp, C.c_int, vp], C.c_int),
        "smalfit_peer_status": ([vp, C.POINTER(C.c_int), vp], C.c_int),
        "smalfit_counters": ([vp, C.POINTER(C.c_int64), vp], C.c_int),
        "smalfit_work_counts": ([vp, C.c_int, C.c_int, C.POINTER(C.c_int64), vp], C.c_int),
    }
    for name, (args, res) in protos.items():
        fn = getattr(lib, name)          # AttributeError here = symbol missing from the .so
        fn.argtypes = args
        fn.restype = res
    if path == LIB_PATH:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "smalfit_abi_version", "smalfit_create", "smalfit_create_ex", "smalfit_status", "smalfit_fused_step", "smalfit_fp32_peak", "smalfit_destroy", "smalfit_last_error", "smalfit_set_targets",
    "smalfit_stage_targets", "smalfit_swap_targets",
    "smalfit_set_visibility", "smalfit_set_masks", "smalfit_set_windows", "smalfit_set_joint_limits", "smalfit_set_focal", "smalfit_set_per_frame_shapes",
    "smalfit_loss_grad", "smalfit_temporal", "smalfit_adam_step", "smalfit_adam_reset", "smalfit_render", "smalfit_vertices", "smalfit_render_color",
    "smalfit_peer_init", "smalfit_peer_connect", "smalfit_peer_allreduce", "smalfit_peer_status",
    "smalfit_counters", "smalfit_work_counts", "smalfit_set_profiling", "smalfit_get_profile",
)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def make_model_desc(c, use_unity_prior: bool = True):
    """Build the smalfit_model_t for a ``model_io.SmalConstants``.  Returns the
    struct and the list of arrays that must stay alive until smalfit_create returns."""
    t = c.tables
    keep = {}

    def F(name, arr):
        keep[name] = _f(arr)
        return keep[name].ctypes.data_as(_f32p)

    def I(name, arr):
        keep[name] = _i(arr)
        return keep[name].ctypes.data_as(_i32p)

    d = ModelDesc()
    d.n_verts = int(c.v_template.shape[0])
    d.n_faces = int(c.faces.shape[0])
    d.v_template = F("v_template", c.v_template)
    d.shapedirs = F("shapedirs", c.shapedirs)
    d.faces = I("faces", c.faces)
    d.parents = I("parents", c.parents)
    d.scale_axis = I("scale_axis", t["scale_axis"])
    d.skin_joint = I("skin_joint", t["skin_joint"])
    d.skin_weight = F("skin_weight", t["skin_weight"])
    for key in ("skinT", "jreg", "mj"):
        setattr(d, f"{key}_ptr", I(f"{key}_ptr", t[f"{key}_ptr"]))
        setattr(d, f"{key}_vert", I(f"{key}_vert", t[f"{key}_vert"]))
        setattr(d, f"{key}_weight", F(f"{key}_weight", t[f"{key}_weight"]))
    for key in ("jregT", "mjT"):
        setattr(d, f"{key}_ptr", I(f"{key}_ptr", t[f"{key}_ptr"]))
        setattr(d, f"{key}_joint", I(f"{key}_joint", t[f"{key}_joint"]))
        setattr(d, f"{key}_weight", F(f"{key}_weight", t[f"{key}_weight"]))
    d.v2f_ptr = I("v2f_ptr", t["v2f_ptr"])
    d.v2f_fc = I("v2f_fc", t["v2f_fc"])
    d.keypoint_joint = I("keypoint_joint", t["keypoint_joint"])
    d.pose_mean = F("pose_mean", c.pose_mean)
    d.pose_prec = F("pose_prec", c.pose_prec)
    d.pose_use = F("pose_use", c.pose_use)
    if use_unity_prior:
        d.shape_dim = 26
        d.shape_mean = F("shape_mean", c.unity_mean)
        d.shape_prec = F("shape_prec", c.unity_prec)
    else:
        d.shape_dim = 20
        d.shape_mean = F("shape_mean", c.cluster_mean)
        d.shape_prec = F("shape_prec", c.cluster_prec)
    return d, keep


class Handle:
    """RAII wrapper of smalfit_t."""

    def __init__(self, constants, device_index: int, max_frames: int, image_size: int, use_unity_prior: bool = True,
                 frame_shard=None, pool_entries_per_frame: int = 0):
        """frame_shard = (lo, hi): this handle only runs the per-frame kernels on frames [lo, hi) (one rank of a
        frame-sharded fit); its workspace and resident targets are sized for those frames only."""
        self.lib = load_library()
        desc, keep = make_model_desc(constants, use_unity_prior)
        h = C.c_void_p()
        opt = Options()
        opt.struct_size = C.sizeof(Options)
        if frame_shard is not None:
            opt.frame_base, opt.frame_capacity = int(frame_shard[0]), int(frame_shard[1]) - int(frame_shard[0])
        opt.pool_entries_per_frame = int(pool_entries_per_frame)
        rc = self.lib.smalfit_create_ex(C.byref(desc), int(device_index), int(max_frames), int(image_size), C.byref(opt), C.byref(h))
        del keep
        if rc != 0:
            raise SmalfitError(f"smalfit_create failed ({rc}): {self.lib.smalfit_last_error(None).decode()}")
        self.h = h
        self.max_frames = max_frames
        self.image_size = image_size
        self.target_set = 0             # which of the two target sets is current (smalfit_swap_targets)

    def check(self, rc: int, what: str):
        if rc != 0:
            raise SmalfitError(f"{what} failed ({rc}): {self.lib.smalfit_last_error(self.h).decode()}")

    def status(self) -> int:
        """Sticky device-side fault bits (STATUS_*), read from host-mapped memory without synchronising."""
        v = C.c_int(0)
        self.lib.smalfit_status(self.h, C.byref(v))
        return int(v.value)

    def raise_on_fault(self):
        st = self.status()
        if st & STATUS_PEER_TIMEOUT:
            raise SmalfitError("a peer rank did not arrive in an all-reduce (fatal: the kernel trapped)")
        if st & STATUS_POOL_OVERFLOW:
            raise SmalfitError("a step needed more (face, tile) entries than the bin pool holds: its silhouette loss and gradient "
                               "were inexact.  Recreate the fitter with a larger pool_entries_per_frame.")

    def stage_targets(self, frame0: int, n: int, sil_ptr, joints_ptr, vis_ptr, from_host: bool, stream_ptr):
        """smalfit_stage_targets: fills the BACK set of targets on the given stream (raw pointers; the caller keeps the
        source tensors alive until that stream has passed the copy)."""
        self.check(self.lib.smalfit_stage_targets(self.h, int(frame0), int(n), sil_ptr, joints_ptr, vis_ptr, 1 if from_host else 0,
                                                  stream_ptr), "smalfit_stage_targets")

    def swap_targets(self) -> int:
        """smalfit_swap_targets: the staged set becomes current for every call enqueued from now on."""
        v = C.c_int(0)
        self.check(self.lib.smalfit_swap_targets(self.h, C.byref(v)), "smalfit_swap_targets")
        self.target_set = int(v.value)
        return self.target_set

    def close(self):
        if getattr(self, "h", None):
            self.lib.smalfit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
