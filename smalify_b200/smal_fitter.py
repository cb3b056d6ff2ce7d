"""SMALFitter: drop-in for ``smal_fitter/smal_fitter.py::SMALFitter`` whose forward
and backward run in libsmalfit's sm_100a kernels.

Same constructor, parameters, attributes and ``forward`` / ``get_temporal`` /
``load_checkpoint`` surface as the reference (smal_fitter.py:25-207); there is
no torch autograd graph, PyTorch3D or CPU path inside: ``forward`` hands the raw
device pointers of the five parameters through the C-ABI, the kernels compute
the loss terms *and* the analytic gradients, and a ``torch.autograd.Function``
shim deposits them when the caller runs ``loss.backward()``.

``FusedFit`` (below) is the opt-in fast path: one call = one epoch of
``optimize_to_joints.py:117-137`` (all windows + temporal + Adam [+ all-reduce])
on flat device buffers, capturable in a CUDA graph.
"""
from __future__ import annotations

import ctypes
import os
import pickle as pkl

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, constants as K, model_io


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _tensors(betas, lbs, glob, joint, trans) -> _cabi.Tensors:
    s = _cabi.Tensors()
    s.betas, s.log_beta_scales, s.global_rotation, s.joint_rotations, s.trans = (
        t.data_ptr() if t is not None else None for t in (betas, lbs, glob, joint, trans))
    return s


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _weights6(weights):
    w = [float(x) for x in weights]
    if len(w) != 6:
        raise ValueError("weights must be (w_j2d, w_reproj, w_betas, w_pose, w_limit, w_splay)")
    return (ctypes.c_float * 6)(*w)


def _contiguous_range(batch_range):
    br = [int(i) for i in batch_range]
    if not br:
        raise ValueError("empty batch_range")
    if br != list(range(br[0], br[0] + len(br))):
        raise ValueError("libsmalfit addresses frames as a contiguous range; got " + repr(br[:8]))
    return br[0], len(br)


class _LossGradFn(torch.autograd.Function):
    """forward: launch the fused kernels (loss terms + analytic gradients into the
    fitter's workspace); backward: scale the stashed gradients by grad_output."""

    @staticmethod
    def forward(ctx, fitter, frame0, n, weights, betas, lbs, glob, joint, trans, focal=None):
        terms = torch.empty(8, device=glob.device, dtype=torch.float32)
        ws = fitter._grad_ws
        lbs_dev = lbs if fitter.use_unity_prior else fitter._zero_logscale
        params = _tensors(betas, lbs_dev, glob, joint, trans)
        grads = _tensors(ws["betas"], ws["log_beta_scales"], ws["global_rotation"], ws["joint_rotations"], ws["trans"])
        h = fitter._handle
        h.check(h.lib.smalfit_loss_grad(h.h, ctypes.byref(params), frame0, n, _weights6(weights), 1,
                                        ctypes.byref(grads), _ptr(terms), _stream(glob.device)), "smalfit_loss_grad")
        ctx.fitter, ctx.frame0, ctx.n = fitter, frame0, n
        # shared-shape gradients are overwritten by the next window: keep this window's copy
        ctx.g_betas = ws["betas"].clone()
        ctx.g_lbs = ws["log_beta_scales"].clone()
        ctx.g_focal = fitter._focal_grad.clone() if focal is not None else None
        ctx.mark_non_differentiable(terms)
        loss = terms[_cabi.L_TOTAL].clone()
        ctx.terms = terms
        return loss, terms

    @staticmethod
    def backward(ctx, grad_loss, _grad_terms):
        f, a, n = ctx.fitter, ctx.frame0, ctx.n
        ws = f._grad_ws
        need = ctx.needs_input_grad          # (fitter, frame0, n, weights, betas, lbs, glob, joint, trans)
        out = [None, None, None, None]
        if f.per_frame_shapes:
            for idx, gsrc in ((4, ctx.g_betas), (5, ctx.g_lbs)):
                if need[idx]:
                    gfull = torch.zeros_like(gsrc)
                    gfull[a:a + n] = grad_loss * gsrc[a:a + n]
                    out.append(gfull)
                else:
                    out.append(None)
        else:
            out.append(grad_loss * ctx.g_betas if need[4] else None)
            if need[5]:
                out.append(grad_loss * ctx.g_lbs if f.use_unity_prior else torch.zeros_like(f.log_beta_scales))
            else:
                out.append(None)
        whole = (a == 0 and n == ws["trans"].shape[0])
        for idx, key in ((6, "global_rotation"), (7, "joint_rotations"), (8, "trans")):
            if need[idx]:
                if whole:                         # one window covers the sequence: no zero fill / slice copy
                    g = grad_loss * ws[key]
                else:
                    g = torch.zeros_like(ws[key])
                    g[a:a + n] = grad_loss * ws[key][a:a + n]
                out.append(g)
            else:
                out.append(None)
        out.append(grad_loss * ctx.g_focal if (len(need) > 9 and need[9] and ctx.g_focal is not None) else None)
        return tuple(out)


class _TemporalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fitter, w_temp, glob, joint, trans):
        dev = glob.device
        terms = torch.empty(3, device=dev, dtype=torch.float32)
        g = {k: torch.zeros_like(v) for k, v in (("g", glob), ("j", joint), ("t", trans))}
        params = _tensors(fitter.betas, fitter._lbs_dev(), glob, joint, trans)
        grads = _tensors(None, None, g["g"], g["j"], g["t"])
        h = fitter._handle
        h.check(h.lib.smalfit_temporal(h.h, ctypes.byref(params), fitter.num_images, float(w_temp),
                                       ctypes.byref(grads), _ptr(terms), _stream(dev)), "smalfit_temporal")
        ctx.g = g
        return terms[0].clone(), terms[1].clone(), terms[2].clone()

    @staticmethod
    def backward(ctx, gj, gg, gt):
        # the kernel returns the gradient of (joint + global + trans); the three terms touch
        # disjoint tensors, so each upstream factor applies to its own tensor
        g = ctx.g
        return None, None, gg * g["g"], gj * g["j"], gt * g["t"]


class SMALFitter(nn.Module):
    """See module docstring.  ``data_batch = (rgb (N,3,S,S), sil (N,1,S,S), joints (N,25,2) (row,col),
    visibility (N,25))`` as produced by the reference loaders (data_loader.py:60-69)."""

    def __init__(self, device, data_batch, batch_size, shape_family, use_unity_prior,
                 constants: model_io.SmalConstants | None = None, data_root: str | None = None,
                 resident_targets: bool = True, per_frame_shapes: bool = False, joint_limits=None, focal=None,
                 frame_shard=None, pool_entries_per_frame: int = 0, binarize_masks: bool = False):
        """Beyond the reference's arguments (smal_fitter.py:26-30):
        frame_shard=(lo, hi): this process fits frames [lo, hi) of the sequence (one rank of a frame-sharded run): the
            library's workspace and resident targets are sized for those frames only; parameters keep their full
            (N, ...) layout.  forward() then takes batch ranges inside [lo, hi).
        pool_entries_per_frame: capacity of the rasteriser's (face, tile) pool (0 = heuristic); an overflow raises.
        binarize_masks: the silhouette targets are stored as uint8 {0, 1}.  The reference takes the L1 distance to
            the float mask (smal_fitter.py:172-173), which is the same thing for the binary masks its loaders
            produce; a mask with other values is rejected unless this flag asks for thresholding at 0.5."""
        super().__init__()
        self.rgb_imgs, self.sil_imgs, self.target_joints, self.target_visibility = data_batch
        self.target_visibility = self.target_visibility.long()
        if self.rgb_imgs is not None:
            assert self.rgb_imgs.max() <= 1.0 and self.rgb_imgs.min() >= 0.0, "RGB Image range is incorrect"

        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.SmalfitError("SMALFitter (B200) needs a CUDA device: the fitting path has no CPU implementation")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_images = int(self.sil_imgs.shape[0])
        self.image_size = int(self.sil_imgs.shape[2])
        self.use_unity_prior = bool(use_unity_prior)
        self.batch_size = int(batch_size)
        self.n_betas = K.N_BETAS
        self.shape_family_list = np.array(shape_family)
        self.resident_targets = resident_targets
        # extension (BASELINE config 4, not in the reference): one shape and one window per frame,
        # i.e. a batch of independent single-image fits
        self.per_frame_shapes = bool(per_frame_shapes)
        if self.per_frame_shapes and not self.use_unity_prior:
            raise NotImplementedError("per_frame_shapes needs the unity prior (trainable log_beta_scales)")

        if constants is None:
            constants = (model_io.load_from_smalify_data(data_root, int(shape_family)) if data_root
                         else model_io.load_asset(shape_family=int(shape_family)))
        self.constants = constants
        dev = self.device
        if self.use_unity_prior:
            mean = torch.from_numpy(constants.unity_mean).float().to(dev)
            self.mean_betas = mean.clone()
            self.betas_prec = torch.from_numpy(constants.unity_prec).float().to(dev)
            if self.per_frame_shapes:
                self.betas = nn.Parameter(mean[:20][None].repeat(self.num_images, 1).contiguous())
                self.log_beta_scales = nn.Parameter(mean[20:][None].repeat(self.num_images, 1).contiguous())
            else:
                self.betas = nn.Parameter(mean[:20].clone())
                self.log_beta_scales = nn.Parameter(mean[20:].clone())
        else:
            self.mean_betas = torch.from_numpy(constants.cluster_mean).float().to(dev)
            self.betas_prec = torch.from_numpy(constants.cluster_prec).float().to(dev)
            self.betas = nn.Parameter(self.mean_betas.clone())
            self.log_beta_scales = nn.Parameter(torch.zeros(self.num_images, 6, device=dev), requires_grad=False)
        self._zero_logscale = torch.zeros(6, device=dev)

        n = self.num_images
        init = torch.tensor(K.GLOBAL_ROT_INIT, dtype=torch.float32, device=dev)
        self.global_rotation = nn.Parameter(init[None].repeat(n, 1))
        self.trans = nn.Parameter(torch.zeros(n, 3, device=dev))
        self.joint_rotations = nn.Parameter(torch.zeros(n, K.N_POSE, 3, device=dev))
        self.global_mask = torch.ones(1, 3, device=dev)
        self.rotation_mask = torch.ones(K.N_POSE, 3, device=dev)

        self.faces = torch.from_numpy(np.asarray(constants.faces).astype(np.int64)).to(dev)
        self.frame_shard = (0, n) if frame_shard is None else (int(frame_shard[0]), int(frame_shard[1]))
        if not (0 <= self.frame_shard[0] < self.frame_shard[1] <= n):
            raise ValueError(f"frame_shard {frame_shard} outside the {n}-frame sequence")
        self._handle = _cabi.Handle(constants, self.device.index, n, self.image_size, self.use_unity_prior,
                                    frame_shard=self.frame_shard, pool_entries_per_frame=pool_entries_per_frame)
        if self.per_frame_shapes:
            self._handle.check(self._handle.lib.smalfit_set_per_frame_shapes(self._handle.h, 1), "smalfit_set_per_frame_shapes")
        # extension: the joint-limit term the reference keeps commented out (smal_fitter.py:77-79,
        # 146-151).  None = off (reference behaviour: w_limit is ignored); True = the reference's own table
        # (constants.joint_limits()); or a (min, max) pair of (34, 3) arrays.
        self.joint_limits = None
        self.set_joint_limits(joint_limits)
        # extension: a focal parameter.  None = the reference's fixed 60-degree camera and no sixth
        # parameter; a float registers `self.focal` (frozen until the caller sets focal.requires_grad = True, like
        # the stage loop does for the other tensors) and dL/dfocal is produced with the other gradients.
        self.focal = None
        self._focal_grad = torch.zeros(1, device=dev)
        if focal is not None:
            self.focal = nn.Parameter(torch.tensor([float(focal)], device=dev), requires_grad=False)
            self._handle.check(self._handle.lib.smalfit_set_focal(self._handle.h, _ptr(self.focal), _ptr(self._focal_grad)),
                               "smalfit_set_focal")
        ns = n if self.per_frame_shapes else 1
        self._grad_ws = {
            "betas": torch.zeros(ns * 20, device=dev).view(self.betas.shape),
            "log_beta_scales": torch.zeros(ns * 6, device=dev).view(6) if ns == 1 else torch.zeros(n, 6, device=dev),
            "global_rotation": torch.zeros(n, 3, device=dev), "joint_rotations": torch.zeros(n, K.N_POSE, 3, device=dev),
            "trans": torch.zeros(n, 3, device=dev)}
        self._windows_for = None
        self._masks_sent = None
        self._vis_sent = None
        # host copies in the layout the library takes (uint8 masks, float32 joints)
        sil_f = self.sil_imgs.reshape(n, self.image_size, self.image_size)
        if not binarize_masks and not bool(((sil_f == 0) | (sil_f == 1)).all()):
            raise ValueError("silhouette targets must be binary {0, 1} masks (the library keeps them as uint8); "
                             "pass binarize_masks=True to threshold soft masks at 0.5")
        self._sil_u8 = (sil_f > 0.5).to(torch.uint8).contiguous()
        self._joints_f32 = self.target_joints.reshape(n, K.N_KEYPOINTS, 2).float().contiguous()
        if self.device.type == "cuda":
            self._sil_u8 = self._sil_u8.pin_memory() if not self._sil_u8.is_cuda else self._sil_u8
            self._joints_f32 = self._joints_f32.pin_memory() if not self._joints_f32.is_cuda else self._joints_f32
        self._upload_targets(self.frame_shard[0], self.frame_shard[1] - self.frame_shard[0])
        self._n_forward = 0

    # ------------------------------------------------------------------
    def check_faults(self):
        """Raises if a kernel reported a sticky fault (bin-pool overflow: inexact silhouette terms; peer timeout).
        Reads a host-mapped word: no synchronisation, so a fault of the step just enqueued may only show on the
        next call."""
        self._handle.raise_on_fault()

    def _lbs_dev(self):
        return self.log_beta_scales if self.use_unity_prior else self._zero_logscale

    def _vis_u8(self, a, n):
        return self.target_visibility[a:a + n].reshape(n, K.N_KEYPOINTS).to(torch.uint8).contiguous()

    def _upload_targets(self, a, n):
        h = self._handle
        vis = self._vis_u8(a, n)
        from_host = 0 if self._sil_u8.is_cuda else 1
        if from_host:
            vis = vis.cpu()
            joints, sil = self._joints_f32[a:a + n], self._sil_u8[a:a + n]
        else:
            vis = vis.to(self.device)
            joints, sil = self._joints_f32[a:a + n], self._sil_u8[a:a + n]
        h.check(h.lib.smalfit_set_targets(h.h, a, n, _ptr(sil), _ptr(joints), _ptr(vis), from_host,
                                          _stream(self.device)), "smalfit_set_targets")
        if from_host:
            torch.cuda.current_stream(self.device).synchronize()     # vis is a temporary host tensor

    def stage_targets(self, sil_u8: torch.Tensor, joints: torch.Tensor, visibility: torch.Tensor, stream: torch.cuda.Stream):
        """Pipelined upload of the NEXT targets of this fitter's frames (not in the reference, which re-sends its
        targets inside every forward, smal_fitter.py:118-120): copies them into the library's second set of target buffers
        on `stream` -- a copy stream, so that the transfer runs under the kernels of the step in flight -- and returns the
        event that marks the copy.  Tensors: (n, S, S) uint8 {0, 1}, (n, 25, 2) float32 (row, col), (n, 25) uint8 for the
        frames of `frame_shard`, pinned host memory or device memory; the caller keeps them alive until the event.
        `swap_targets()` then makes them the targets of every later call."""
        a, b = self.frame_shard
        n = b - a
        if tuple(sil_u8.shape) != (n, self.image_size, self.image_size) or sil_u8.dtype != torch.uint8:
            raise ValueError(f"stage_targets: masks must be uint8 of shape {(n, self.image_size, self.image_size)}")
        if tuple(joints.shape) != (n, K.N_KEYPOINTS, 2) or joints.dtype != torch.float32:
            raise ValueError(f"stage_targets: joints must be float32 of shape {(n, K.N_KEYPOINTS, 2)}")
        if tuple(visibility.shape) != (n, K.N_KEYPOINTS) or visibility.dtype != torch.uint8:
            raise ValueError(f"stage_targets: visibility must be uint8 of shape {(n, K.N_KEYPOINTS)}")
        devs = {t.is_cuda for t in (sil_u8, joints, visibility)}
        if len(devs) != 1:
            raise ValueError("stage_targets: all three tensors on the host (pinned) or all on the device")
        from_host = not sil_u8.is_cuda
        if from_host and not all(t.is_pinned() for t in (sil_u8, joints, visibility)):
            raise ValueError("stage_targets: host tensors must be pinned (the copy is asynchronous)")
        if not all(t.is_contiguous() for t in (sil_u8, joints, visibility)):
            raise ValueError("stage_targets: tensors must be contiguous")
        self._handle.stage_targets(a, n, _ptr(sil_u8), _ptr(joints), _ptr(visibility), from_host, ctypes.c_void_p(stream.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def swap_targets(self, staged: torch.cuda.Event | None = None) -> int:
        """Makes the staged targets current (the current stream first waits for `staged`, the event `stage_targets`
        returned).  Returns the index (0 / 1) of the set now in use; `FusedFit` keeps one CUDA graph per set.  The
        visibility rows of the new set are the staged ones: `target_visibility` is not re-sent over them."""
        if staged is not None:
            torch.cuda.current_stream(self.device).wait_event(staged)
        idx = self._handle.swap_targets()
        # the staged rows stand until `target_visibility` changes again (the attributes sil_imgs / target_joints /
        # target_visibility keep describing the targets the fitter was constructed with)
        self._vis_sent = (self._vis_key(), {(self.frame_shard[0], self.frame_shard[1] - self.frame_shard[0])})
        return idx

    def _vis_key(self):
        tv = self.target_visibility
        # (host tensor, as the loaders produce: its content is the key -- a rebound tensor can reuse a freed address)
        return hash(tv.numpy().tobytes()) if tv.device.type == "cpu" else (tv.data_ptr(), tv._version)

    def _sync_visibility(self, a, n):
        """The stage loop rewrites target_visibility in place or rebinds it (optimize_to_joints.py:98-110): the rows
        are sent again only when the tensor or its version counter changed since they were last sent."""
        key = self._vis_key()
        if self._vis_sent is None or self._vis_sent[0] != key:
            self._vis_sent = (key, set())
        if any(sa <= a and a + n <= sa + sn for sa, sn in self._vis_sent[1]):
            return
        vis = self._vis_u8(a, n).to(self.device, non_blocking=False)
        h = self._handle
        h.check(h.lib.smalfit_set_visibility(h.h, a, n, _ptr(vis), 0, _stream(self.device)), "smalfit_set_visibility")
        self._vis_keepalive = vis
        self._vis_sent[1].add((a, n))

    def _sync_masks(self):
        key = (self.global_mask._version, self.global_mask.data_ptr(), self.rotation_mask._version, self.rotation_mask.data_ptr())
        if key == self._masks_sent:
            return
        gm = np.ascontiguousarray(self.global_mask.detach().cpu().numpy().reshape(3), dtype=np.float32)
        rm = np.ascontiguousarray(self.rotation_mask.detach().cpu().numpy().reshape(K.N_POSE * 3), dtype=np.float32)
        h = self._handle
        fp = ctypes.POINTER(ctypes.c_float)
        h.check(h.lib.smalfit_set_masks(h.h, gm.ctypes.data_as(fp), rm.ctypes.data_as(fp)), "smalfit_set_masks")
        self._masks_sent = key

    def set_windows(self, frames_per_window):
        """frames_per_window[i] = size of the window frame i is optimised in (the B of the
        reference's mean() reductions)."""
        arr = np.ascontiguousarray(frames_per_window, dtype=np.int32)
        h = self._handle
        h.check(h.lib.smalfit_set_windows(h.h, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(arr)),
                "smalfit_set_windows")

    def _set_window_for(self, a, n):
        if self.per_frame_shapes:
            if self._windows_for is None:
                self._windows_for = np.ones(self.num_images, dtype=np.int32)
                self.set_windows(self._windows_for)
            return
        if self._windows_for is None:
            self._windows_for = np.full(self.num_images, self.num_images, dtype=np.int32)
            self._windows_for[:] = -1
        if not np.all(self._windows_for[a:a + n] == n):
            self._windows_for[a:a + n] = n
            arr = np.where(self._windows_for > 0, self._windows_for, self.num_images).astype(np.int32)
            self.set_windows(arr)

    # ------------------------------------------------------------------
    def forward(self, batch_range, weights, stage_id):
        """SMALFitter.forward (smal_fitter.py:107-175): returns (loss, objs)."""
        a, n = _contiguous_range(batch_range)
        self._n_forward += 1
        self._sync_masks()
        self._set_window_for(a, n)
        if self.resident_targets:
            self._sync_visibility(a, n)
        else:
            self._upload_targets(a, n)         # the per-call H2D copy of smal_fitter.py:118-120
        loss, terms = _LossGradFn.apply(self, a, n, [float(w) for w in weights], self.betas, self.log_beta_scales,
                                        self.global_rotation, self.joint_rotations, self.trans, self.focal)
        w_j2d, w_reproj, w_betas, w_pose, w_limit, w_splay = [float(w) for w in weights]
        objs = {}
        if w_j2d > 0:
            objs["joint"] = terms[_cabi.L_JOINT]
        if w_limit > 0 and self.joint_limits is not None:
            objs["limit"] = terms[_cabi.L_LIMIT]
        if w_pose > 0:
            objs["pose"] = terms[_cabi.L_POSE]
        if w_splay > 0:
            objs["splay"] = terms[_cabi.L_SPLAY]
        if w_betas > 0:
            objs["betas"] = terms[_cabi.L_BETAS]
        if w_reproj > 0:
            objs["sil_reproj"] = terms[_cabi.L_SIL]
        return loss, objs

    def set_joint_limits(self, limits):
        h = self._handle
        fp = ctypes.POINTER(ctypes.c_float)
        if limits is None or limits is False:
            h.check(h.lib.smalfit_set_joint_limits(h.h, None, None), "smalfit_set_joint_limits")
            self.joint_limits = None
            return
        lo, hi = K.joint_limits() if limits is True else limits
        lo = np.ascontiguousarray(np.asarray(lo, np.float32).reshape(K.N_POSE, 3))
        hi = np.ascontiguousarray(np.asarray(hi, np.float32).reshape(K.N_POSE, 3))
        h.check(h.lib.smalfit_set_joint_limits(h.h, lo.ctypes.data_as(fp), hi.ctypes.data_as(fp)), "smalfit_set_joint_limits")
        self.joint_limits = (lo, hi)

    def get_temporal(self, w_temp):
        """smal_fitter.py:177-190: returns (joint_loss, global_loss, trans_loss)."""
        self._sync_masks()
        return _TemporalFn.apply(self, float(w_temp), self.global_rotation, self.joint_rotations, self.trans)

    # ------------------------------------------------------------------
    @torch.no_grad()
    def render(self, batch_range=None):
        """Soft silhouettes (B,1,S,S) and projected keypoints (B,25,2) of the current
        parameters (Renderer.forward's first two outputs, p3d_renderer.py:61-74)."""
        a, n = _contiguous_range(batch_range if batch_range is not None else range(*self.frame_shard))
        self._sync_masks()
        S = self.image_size
        sil = torch.empty(n, 1, S, S, device=self.device)
        kp = torch.empty(n, K.N_KEYPOINTS, 2, device=self.device)
        params = _tensors(self.betas, self._lbs_dev(), self.global_rotation, self.joint_rotations, self.trans)
        h = self._handle
        h.check(h.lib.smalfit_render(h.h, ctypes.byref(params), a, n, _ptr(sil), _ptr(kp), _stream(self.device)),
                "smalfit_render")
        return sil, kp

    @torch.no_grad()
    def vertices(self, batch_range=None):
        a, n = _contiguous_range(batch_range if batch_range is not None else range(*self.frame_shard))
        self._sync_masks()
        v = torch.empty(n, self.constants.v_template.shape[0], 3, device=self.device)
        params = _tensors(self.betas, self._lbs_dev(), self.global_rotation, self.joint_rotations, self.trans)
        h = self._handle
        h.check(h.lib.smalfit_vertices(h.h, ctypes.byref(params), a, n, _ptr(v), _stream(self.device)), "smalfit_vertices")
        return v

    # ---- visualisation (row 8f-3; host-side glue around smalfit_render_color) ----------------------
    @torch.no_grad()
    def render_color(self, verts: torch.Tensor, color=None) -> torch.Tensor:
        """Hard Phong rendering (B,3,S,S) of world-space vertices (B,V,3): the reference's colour renderer
        (p3d_renderer.py:41-59,70-72).  B <= num_images."""
        from .visualization import MESH_COLOR
        v = verts.detach().to(self.device, torch.float32).contiguous()
        n, S = int(v.shape[0]), self.image_size
        out = torch.empty(n, 3, S, S, device=self.device)
        col = (ctypes.c_float * 3)(*(color if color is not None else MESH_COLOR))
        h = self._handle
        h.check(h.lib.smalfit_render_color(h.h, _ptr(v), n, col, _ptr(out), _stream(self.device)), "smalfit_render_color")
        return out

    @torch.no_grad()
    def model_joints(self, batch_range=None) -> torch.Tensor:
        """The 41 model joints (B,41,3) of the current parameters, trans included (smal_torch.py:171-184)."""
        if not hasattr(self, "_mj_dense"):
            t = self.constants.tables
            V = self.constants.v_template.shape[0]
            M = np.zeros((K.N_MODEL_JOINTS, V), np.float32)
            for j in range(K.N_MODEL_JOINTS):
                sl = slice(int(t["mj_ptr"][j]), int(t["mj_ptr"][j + 1]))
                M[j, t["mj_vert"][sl]] = t["mj_weight"][sl]
            self._mj_dense = torch.from_numpy(M).to(self.device)
        return torch.einsum("jv,bvk->bjk", self._mj_dense, self.vertices(batch_range))     # rows sum to 1: trans carries over

    @torch.no_grad()
    def project_points(self, points: torch.Tensor) -> torch.Tensor:
        """(B,J,3) world points -> (B,J,2) (row, col) pixels: transform_points_screen as used at p3d_renderer.py:67-68."""
        f = float(self.focal) if self.focal is not None else K.CAMERA_FOCAL
        zv = K.CAMERA_DISTANCE - points[..., 2]
        xn, yn = -f * points[..., 0] / zv, f * points[..., 1] / zv
        half = (self.image_size - 1) / 2.0
        return torch.stack([half * (1.0 - yn), half * (1.0 - xn)], dim=-1)

    def generate_visualization(self, image_exporter):
        """smal_fitter.py:209-272: one collage + parameter pickle + mesh per frame through the exporter."""
        from .visualization import collage
        for j in range(0, self.num_images, self.batch_size):
            br = list(range(j, min(self.num_images, j + self.batch_size)))
            rows = collage(self, br)
            verts = self.vertices(br).cpu()
            for b, gid in enumerate(br):
                img = (np.transpose(rows[b].numpy(), (1, 2, 0)) * 255.0).astype(np.uint8)
                image_exporter.export(img, b, gid, self.export_parameters(gid), verts, np.asarray(self.constants.faces))

    def set_profiling(self, enable, count_pairs: bool = False):
        """Per-phase CUDA events on every loss_grad call; count_pairs additionally makes the backward count the
        pairs work_counts() reports (slows it down: not for timed passes)."""
        h = self._handle
        h.check(h.lib.smalfit_set_profiling(h.h, 2 if (enable and count_pairs) else int(bool(enable))), "smalfit_set_profiling")

    def profile(self):
        """Per-phase device milliseconds of the last loss_grad call (see smalfit_get_profile)."""
        arr = (ctypes.c_float * 8)()
        h = self._handle
        h.check(h.lib.smalfit_get_profile(h.h, arr), "smalfit_get_profile")
        keys = ("pose_forward", "face_rects", "raster_forward", "raster_backward", "frame_backward", "shape_backward", "total")
        return {k: float(arr[i]) for i, k in enumerate(keys)}

    def counters(self):
        arr = (ctypes.c_int64 * 4)()
        h = self._handle
        h.check(h.lib.smalfit_counters(h.h, arr, _stream(self.device)), "smalfit_counters")
        return dict(capped_pixels=arr[0], spilled_pixels=arr[1], dropped_bin_entries=arr[2], launches=arr[3])

    def work_counts(self, frame0=None, n=None):
        """(pixel, face) pairs and (face, tile) entries of the last rasterised pass, and -- accumulated by the
        backward while profiling is on -- the pairs in pixels that carry a gradient / that contribute to it
        (diagnostic, synchronises)."""
        arr = (ctypes.c_int64 * 4)()
        h = self._handle
        frame0 = self.frame_shard[0] if frame0 is None else frame0
        n = self.frame_shard[1] - frame0 if n is None else n
        h.check(h.lib.smalfit_work_counts(h.h, frame0, n, arr, _stream(self.device)), "smalfit_work_counts")
        return dict(pairs=arr[0], tile_entries=arr[1], live_pairs=arr[2], used_pairs=arr[3])

    def fp32_peak(self):
        """Measured FP32 FMA ceiling of this GPU in TFLOP/s: scalar FFMA and packed FFMA2 (roofline denominator)."""
        arr = (ctypes.c_float * 2)()
        h = self._handle
        h.check(h.lib.smalfit_fp32_peak(h.h, arr, _stream(self.device)), "smalfit_fp32_peak")
        return dict(ffma=float(arr[0]), ffma2=float(arr[1]))

    # ------------------------------------------------------------------
    def export_parameters(self, frame_id):
        """The per-frame dict ImageExporter pickles (smal_fitter.py:213-219,268): this frame's rotations and
        translation, and the shape the frame is rendered with (the shared one, or the frame's own when every frame
        has its own shape): betas (20,), log_betascale (6,)."""
        with torch.no_grad():
            betas = self.betas[frame_id] if self.per_frame_shapes else self.betas
            if self.per_frame_shapes or not self.use_unity_prior:
                scales = self.log_beta_scales[frame_id]
            else:
                scales = self.log_beta_scales
            return {
                "global_rotation": (self.global_rotation[frame_id] * self.global_mask[0]).cpu().numpy(),
                "joint_rotations": (self.joint_rotations[frame_id] * self.rotation_mask).cpu().numpy(),
                "betas": betas.detach().cpu().numpy(),
                "log_betascale": scales.detach().cpu().numpy(),
                "trans": self.trans[frame_id].detach().cpu().numpy(),
            }

    def load_checkpoint(self, checkpoint_path, epoch):
        """smal_fitter.py:192-207: per-frame pickles; the shared betas / scales become the mean over frames (with one
        shape per frame every frame keeps its own).  Values are copied into the existing parameter storage, so the
        library handle and any FusedFit views stay valid."""
        beta_list, scale_list = [], []
        with torch.no_grad():
            for frame_id in range(self.num_images):
                param_file = os.path.join(checkpoint_path, "{0:04}".format(frame_id), "{0}.pkl".format(epoch))
                with open(param_file, "rb") as f:
                    p = pkl.load(f)
                self.global_rotation[frame_id] = torch.from_numpy(p["global_rotation"]).float().to(self.device)
                self.joint_rotations[frame_id] = torch.from_numpy(p["joint_rotations"]).float().to(self.device).view(K.N_POSE, 3)
                self.trans[frame_id] = torch.from_numpy(p["trans"]).float().to(self.device)
                beta_list.append(np.asarray(p["betas"]).reshape(-1)[:self.n_betas])
                scale_list.append(np.asarray(p["log_betascale"]).reshape(-1)[:6])
            betas = torch.from_numpy(np.stack(beta_list)).float().to(self.device)
            scales = torch.from_numpy(np.stack(scale_list)).float().to(self.device)
            if self.per_frame_shapes:
                self.betas.copy_(betas)
                self.log_beta_scales.copy_(scales)
            else:
                self.betas.copy_(betas.mean(0))
                self.log_beta_scales.copy_(scales.mean(0) if self.use_unity_prior else scales)


class FusedFit:
    """One call = one epoch of optimize_to_joints.py:117-137 on flat device buffers, as ONE library call
    (smalfit_fused_step): loss + gradients of every window with the temporal term folded in, then a single tail
    kernel that [exchanges the gradient with the peer ranks over NVLink and] applies Adam.  Parameters stay the
    fitter's nn.Parameters (re-pointed at slices of one flat buffer).

    collective (frames sharded over `process_group` only):
      "peer"  the exchange lives inside the tail kernel (peer memory, rank-ordered sums, bit-identical replicas);
              needs equal contiguous shards, rank r owning frames [r n, (r + 1) n)
      "nccl"  the unfused sequence: loss_grad, one torch.distributed.all_reduce of the flat gradient, temporal, Adam
    With one shape per frame (`per_frame_shapes`, BASELINE config 4) nothing is shared between frames: no
    collective at all, every rank steps its own frames only."""

    NAMES = ("betas", "log_beta_scales", "global_rotation", "joint_rotations", "trans")

    def __init__(self, fitter: SMALFitter, window_size: int | None = None, frame_shard=None, process_group=None,
                 collective: str = "peer"):
        self.f = fitter
        n = fitter.num_images
        dev = fitter.device
        if not fitter.use_unity_prior:
            raise NotImplementedError("FusedFit supports the unity-prior (shared log_beta_scales) configuration")
        if collective not in ("nccl", "peer"):
            raise ValueError("collective must be 'nccl' or 'peer'")
        ns = n if fitter.per_frame_shapes else 1
        self.sizes = (ns * 20, ns * 6, n * 3, n * K.N_POSE * 3, n * 3)
        total = sum(self.sizes)
        nt = _cabi.N_TERMS_FUSED
        self.flat_p = torch.empty(total, device=dev)
        self.flat_g = torch.zeros(total + nt, device=dev)       # + the loss terms: one all-reduce covers both (nccl path)
        self.flat_m = torch.zeros(total, device=dev)
        self.flat_v = torch.zeros(total, device=dev)
        shapes = (tuple(fitter.betas.shape), tuple(fitter.log_beta_scales.shape), (n, 3), (n, K.N_POSE, 3), (n, 3))
        off = 0
        self.views = {}
        with torch.no_grad():
            for name, size, shape in zip(self.NAMES, self.sizes, shapes):
                par = getattr(fitter, name)
                self.flat_p[off:off + size] = par.detach().reshape(-1)
                par.data = self.flat_p[off:off + size].view(shape)
                self.views[name] = tuple(buf[off:off + size] for buf in (self.flat_p, self.flat_g, self.flat_m, self.flat_v))
                off += size
        self.window = 1 if fitter.per_frame_shapes else (window_size or n)
        wins = np.empty(n, dtype=np.int32)
        for j in range(0, n, self.window):
            wins[j:j + self.window] = min(self.window, n - j)
        fitter.set_windows(wins)
        fitter._windows_for = wins.copy()
        self.n_windows = (n + self.window - 1) // self.window
        self.shard = tuple(frame_shard) if frame_shard is not None else tuple(fitter.frame_shard)
        if not (fitter.frame_shard[0] <= self.shard[0] and self.shard[1] <= fitter.frame_shard[1]):
            raise ValueError(f"frame_shard {self.shard} is outside the frames the fitter holds {fitter.frame_shard}")
        self.group = process_group
        self.sharded = process_group is not None and not fitter.per_frame_shapes
        self.terms = self.flat_g[total:total + nt]
        self.temporal_terms = self.terms[8:11]
        self.step_count = 0
        self._graph = None
        self._graph_key = None
        self._graphs = {}
        self._warmed = False
        self.peer_error = None
        self.collective = None
        if self.sharded:
            self.collective = "nccl"
            if collective == "peer":
                self._connect_peers(total + nt)
        self.fused_tail = (not self.sharded) or self.collective == "peer"
        if not self.fused_tail:
            self.temporal_terms = torch.zeros(3, device=dev)

    def _connect_peers(self, n_floats: int):
        """Sets up the peer-memory exchange on every rank, or on none: the ranks agree (MIN over a success
        flag) after each step, so that a GPU without P2P / IPC access -- or shards the tail kernel cannot address --
        leaves all of them on NCCL."""
        import torch.distributed as dist
        h, dev, group = self.f._handle, self.f.device, self.group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        on_gpu = dist.get_backend(group) == "nccl"
        cdev = dev if on_gpu else "cpu"

        def all_ok(ok: bool) -> bool:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=cdev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())

        n = self.f.num_images
        per = self.shard[1] - self.shard[0]
        if not all_ok(per * world == n and self.shard[0] == rank * per and world <= 8):
            self.peer_error = "unequal / non-contiguous shards (or more than 8 ranks)"
            return
        mine_c = (ctypes.c_ubyte * 64)()
        rc = h.lib.smalfit_peer_init(h.h, rank, world, int(n_floats), mine_c)
        if not all_ok(rc == 0):
            self.peer_error = "smalfit_peer_init failed on some rank"
            return
        mine = torch.tensor(list(mine_c), dtype=torch.uint8, device=cdev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        blob = torch.cat(gathered).cpu().numpy().tobytes()
        rc = h.lib.smalfit_peer_connect(h.h, (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob))
        if not all_ok(rc == 0):                   # (also the barrier: every rank has mapped every buffer before the first store)
            self.peer_error = "smalfit_peer_connect failed on some rank: " + h.lib.smalfit_last_error(h.h).decode()
            return
        self.collective = "peer"

    def peer_timed_out(self) -> bool:
        """True when a peer failed to arrive in some exchange (fatal: the kernel trapped)."""
        return bool(self.f._handle.status() & _cabi.STATUS_PEER_TIMEOUT)

    def _t(self, idx):
        return _tensors(*(self.views[k][idx] for k in self.NAMES))

    def reset_optimizer(self):
        """Fresh Adam state, as a new torch.optim.Adam per stage (optimize_to_joints.py:96)."""
        self.flat_m.zero_()
        self.flat_v.zero_()
        self.step_count = 0
        h = self.f._handle
        h.check(h.lib.smalfit_adam_reset(h.h, _stream(self.f.device)), "smalfit_adam_reset")
        self._graph, self._graphs = None, {}

    def _enqueue(self, weights, w_temp, lr, train, device_step: bool):
        f = self.f
        h = f._handle
        st = _stream(f.device)
        a, b = self.shard
        params, grads = self._t(0), self._t(1)
        rank0 = (not self.sharded) or torch.distributed.get_rank(self.group) == 0
        tr = (ctypes.c_int32 * 5)(*[int(x) for x in train])
        m, v = self._t(2), self._t(3)
        if self.fused_tail:
            h.check(h.lib.smalfit_fused_step(h.h, ctypes.byref(params), ctypes.byref(grads), ctypes.byref(m), ctypes.byref(v),
                                             a, b - a, f.num_images, _weights6(weights), float(w_temp),
                                             self.n_windows if rank0 else 0, tr, float(lr), K.ADAM_BETAS[0], K.ADAM_BETAS[1],
                                             K.ADAM_EPS, _ptr(self.terms), st), "smalfit_fused_step")
            return
        # unfused sequence with NCCL: frames this rank does not own must contribute zeros to the sum
        self.flat_g.zero_()
        h.check(h.lib.smalfit_loss_grad(h.h, ctypes.byref(params), a, b - a, _weights6(weights),
                                        self.n_windows if rank0 else 0, ctypes.byref(grads), _ptr(self.terms), st),
                "smalfit_loss_grad")
        torch.distributed.all_reduce(self.flat_g, group=self.group)      # gradients + loss terms
        h.check(h.lib.smalfit_temporal(h.h, ctypes.byref(params), f.num_images, float(w_temp), ctypes.byref(grads),
                                       _ptr(self.temporal_terms), st), "smalfit_temporal")
        h.check(h.lib.smalfit_adam_step(h.h, ctypes.byref(params), ctypes.byref(grads), ctypes.byref(m), ctypes.byref(v),
                                        f.num_images, tr, float(lr), K.ADAM_BETAS[0], K.ADAM_BETAS[1], K.ADAM_EPS,
                                        0 if device_step else self.step_count, st), "smalfit_adam_step")

    def step(self, weights, w_temp, lr, train=(1, 1, 1, 1, 1), use_graph: bool = False):
        """Returns nothing; self.terms / self.temporal_terms hold the loss terms on the device."""
        if self.f.focal is not None and self.f.focal.requires_grad:
            raise NotImplementedError("FusedFit keeps the focal parameter fixed; train it through SMALFitter.forward/backward")
        self.step_count += 1
        f = self.f
        f._sync_masks()
        f.check_faults()                  # host-mapped word: no synchronisation (also covers the graph replays below)
        if not use_graph or not self._warmed:
            # eager launch (also the first call: loads every kernel before a capture)
            self._enqueue(weights, w_temp, lr, train, device_step=True)
            self._warmed = True
            return
        key = (tuple(float(w) for w in weights), float(w_temp), float(lr), tuple(int(t) for t in train))
        tset = f._handle.target_set                # a graph replays the target buffers it was captured with
        if self._graph_key != key:
            self._graphs, self._graph_key = {}, key
        self._graph = self._graphs.get(tset)
        if self._graph is None:
            # warm-up on a side stream, then capture
            s = torch.cuda.Stream(device=f.device)
            s.wait_stream(torch.cuda.current_stream(f.device))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    self._enqueue(weights, w_temp, lr, train, device_step=True)
            torch.cuda.current_stream(f.device).wait_stream(s)
            self._graph = self._graphs[tset] = g
        self._graph.replay()

    def total_loss(self) -> torch.Tensor:
        if self.fused_tail:
            return self.terms[_cabi.L_TOTAL]              # includes the temporal term
        return self.terms[_cabi.L_TOTAL] + self.temporal_terms.sum()
