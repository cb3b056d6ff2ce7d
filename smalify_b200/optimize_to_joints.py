"""The 4-stage fitting loop of ``smal_fitter/optimize_to_joints.py:90-137``.

``fit_sequence(model, ...)`` reproduces the reference's loop (new Adam per stage,
stage-0 freezing and torso-only visibility, windowed accumulation, temporal term)
on a ``SMALFitter``; ``fused=True`` runs every epoch as one ``FusedFit.step``.
``image_exporter`` (data_io.ResultExporter) turns on the reference's visualisation hook: a collage, the
parameter pickle and the mesh of every frame each ``vis_frequency`` epochs and once more at the end, under
the names the reference uses (optimize_to_joints.py:113-115,139-144).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import constants as K
from .smal_fitter import FusedFit, SMALFitter


def stage_visibility(visibility: torch.Tensor, stage_id: int) -> torch.Tensor:
    """optimize_to_joints.py:98-110: torso joints only during stage 0."""
    if stage_id == 0:
        v = torch.zeros_like(visibility)
        v[:, list(K.TORSO_JOINTS)] = visibility[:, list(K.TORSO_JOINTS)]
        return v
    return visibility.clone()


def fit_sequence(model: SMALFitter, schedule=K.STAGE_SCHEDULE, window_size: int = K.WINDOW_SIZE,
                 allow_limb_scaling: bool = True, fused: bool = False, use_graph: bool = False,
                 iters_override=None, callback=None, process_group=None, frame_shard=None,
                 image_exporter=None, vis_frequency: int = K.VIS_FREQUENCY):
    """Runs the stage loop in place on ``model``.  Returns the list of per-stage final losses."""

    def visualise(stage_id, epoch_id):
        if image_exporter is not None and epoch_id % vis_frequency == 0:
            image_exporter.stage_id, image_exporter.epoch_name = stage_id, str(epoch_id)
            model.generate_visualization(image_exporter)

    data_visibility = model.target_visibility.clone()
    n = model.num_images
    fused_loop = FusedFit(model, window_size, frame_shard=frame_shard, process_group=process_group) if fused else None
    finals = []
    for stage_id, row in enumerate(schedule):
        opt_weight, w_temp, epochs, lr = row[:6], row[6], int(row[7]), row[8]
        if iters_override is not None:
            epochs = int(iters_override[stage_id])
        if stage_id == 0:
            model.joint_rotations.requires_grad = False
            model.betas.requires_grad = False
            model.log_beta_scales.requires_grad = False
            model.target_visibility = stage_visibility(data_visibility, 0)
        else:
            model.joint_rotations.requires_grad = True
            model.betas.requires_grad = True
            if allow_limb_scaling and model.use_unity_prior:
                model.log_beta_scales.requires_grad = True
            model.target_visibility = data_visibility.clone()
        last = None
        if fused:
            fused_loop.reset_optimizer()
            model._sync_visibility(model.frame_shard[0], model.frame_shard[1] - model.frame_shard[0])
            train = (int(model.betas.requires_grad), int(model.log_beta_scales.requires_grad), 1,
                     int(model.joint_rotations.requires_grad), 1)
            for epoch_id in range(epochs):
                fused_loop.step(opt_weight, w_temp, lr, train, use_graph=use_graph)
                if callback is not None:
                    callback(stage_id, epoch_id, fused_loop.total_loss())
                visualise(stage_id, epoch_id)
            last = fused_loop.total_loss() if epochs else None
        else:
            optimizer = torch.optim.Adam(model.parameters(), lr=lr, betas=K.ADAM_BETAS)
            for epoch_id in range(epochs):
                acc_loss = 0
                optimizer.zero_grad()
                for j in range(0, n, window_size):
                    batch_range = list(range(j, min(n, j + window_size)))
                    loss, _ = model(batch_range, opt_weight, stage_id)
                    acc_loss = acc_loss + loss.mean()
                joint_loss, global_loss, trans_loss = model.get_temporal(w_temp)
                acc_loss = acc_loss + joint_loss + global_loss + trans_loss
                acc_loss.backward()
                optimizer.step()
                last = acc_loss.detach()
                if callback is not None:
                    callback(stage_id, epoch_id, last)
                visualise(stage_id, epoch_id)
        finals.append(None if last is None else float(last))
    if image_exporter is not None:                   # "Final stage" export (optimize_to_joints.py:142-144)
        image_exporter.stage_id, image_exporter.epoch_name = 10, "0"
        model.generate_visualization(image_exporter)
    return finals


# ------------------------------------------------------------------------------------------------
# Script level: optimize_to_joints.main (:55-144) and generate_video.main (generate_video.py:38-74)
# ------------------------------------------------------------------------------------------------
class RunConfig:
    """The run settings of the reference's config.py (:5-31), as an object instead of a module."""

    def __init__(self, **kw):
        import time
        self.data_path = "data"
        self.BADJA_PATH = "data/BADJA"
        self.STANFORD_EXTRA_PATH = "data/StanfordExtra"
        self.OUTPUT_DIR = "checkpoints/{0}".format(time.strftime("%Y%m%d-%H%M%S"))
        self.CROP_SIZE = K.CROP_SIZE
        self.VIS_FREQUENCY = K.VIS_FREQUENCY
        self.FORCE_SMAL_PRIOR = False
        self.ALLOW_LIMB_SCALING = True
        self.SHAPE_FAMILY = 1
        self.SEQUENCE_OR_IMAGE_NAME = "badja:rs_dog"
        self.IMAGE_RANGE = range(0, 1)
        self.WINDOW_SIZE = K.WINDOW_SIZE
        self.CHECKPOINT_NAME = ""
        self.EPOCH_NAME = "st10_ep0"
        self.OPT_SCHEDULE = K.STAGE_SCHEDULE
        self.FUSED = True               # one FusedFit.step (CUDA graph) per epoch instead of the autograd shim
        self.EXPORT = True
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown setting " + k)
            setattr(self, k, v)


def load_run_data(cfg: RunConfig):
    from . import data_io
    dataset, name = cfg.SEQUENCE_OR_IMAGE_NAME.split(":")
    if dataset == "badja":
        return data_io.load_badja_sequence(cfg.BADJA_PATH, name, cfg.CROP_SIZE, image_range=cfg.IMAGE_RANGE)
    return data_io.load_stanford_sequence(cfg.STANFORD_EXTRA_PATH, name, cfg.CROP_SIZE)


def _make_fitter(cfg: RunConfig, data, constants=None):
    assert cfg.SHAPE_FAMILY >= 0, "Shape family should be greater than 0"
    use_unity_prior = cfg.SHAPE_FAMILY == 1 and not cfg.FORCE_SMAL_PRIOR
    if not use_unity_prior and cfg.ALLOW_LIMB_SCALING:
        print("WARNING: Limb scaling is only recommended for the new Unity prior.")
        cfg.ALLOW_LIMB_SCALING = False
    if constants is None and cfg.data_path and os.path.isdir(os.path.join(cfg.data_path, "SMALST")):
        return SMALFitter("cuda", data, cfg.WINDOW_SIZE, cfg.SHAPE_FAMILY, use_unity_prior, data_root=cfg.data_path)
    return SMALFitter("cuda", data, cfg.WINDOW_SIZE, cfg.SHAPE_FAMILY, use_unity_prior, constants=constants)


def main(cfg: RunConfig | None = None, constants=None, data=None, iters_override=None):
    """optimize_to_joints.main: load the sequence / image, fit it with the 4-stage schedule, export a collage +
    parameters + mesh per frame every VIS_FREQUENCY epochs and at the end (st10_ep0).  Returns (fitter, finals)."""
    from . import data_io
    cfg = cfg or RunConfig()
    os.makedirs(cfg.OUTPUT_DIR, exist_ok=True)
    data, filenames = data if data is not None else load_run_data(cfg)
    print("Dataset size: {0}".format(len(filenames)))
    model = _make_fitter(cfg, data, constants)
    exporter = data_io.ResultExporter(cfg.OUTPUT_DIR, filenames) if cfg.EXPORT else None
    fused = bool(cfg.FUSED and model.use_unity_prior)
    finals = fit_sequence(model, cfg.OPT_SCHEDULE, cfg.WINDOW_SIZE, allow_limb_scaling=cfg.ALLOW_LIMB_SCALING, fused=fused,
                          use_graph=fused, iters_override=iters_override, image_exporter=exporter, vis_frequency=cfg.VIS_FREQUENCY)
    return model, finals


class FrameExporter:
    """generate_video.py:26-36: <output_dir>/<%04d>.png + .pkl per frame."""

    def __init__(self, output_dir):
        os.makedirs(output_dir, exist_ok=True)
        self.output_dir = output_dir

    def export(self, collage_np, batch_id, global_id, img_parameters, vertices, faces):
        import pickle as pkl
        import cv2
        stem = os.path.join(self.output_dir, "{0:04}".format(global_id))
        cv2.imwrite(stem + ".png", np.ascontiguousarray(collage_np[:, :, ::-1]))
        with open(stem + ".pkl", "wb") as f:
            pkl.dump(img_parameters, f)


def generate_video(cfg: RunConfig, constants=None, data=None, checkpoints_root: str = "checkpoints", export_root: str = "exported"):
    """generate_video.main: reload a finished fit from its per-frame pickles and write one collage per frame
    (the ffmpeg call that joins them is the user's, as in the reference)."""
    data, filenames = data if data is not None else load_run_data(cfg)
    model = _make_fitter(cfg, data, constants)
    model.load_checkpoint(os.path.join(checkpoints_root, cfg.CHECKPOINT_NAME), cfg.EPOCH_NAME)
    out = os.path.join(export_root, cfg.CHECKPOINT_NAME, cfg.EPOCH_NAME)
    model.generate_visualization(FrameExporter(out))
    return out


if __name__ == "__main__":
    main()
