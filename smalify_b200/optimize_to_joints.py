"""The 4-stage fitting loop of ``smal_fitter/optimize_to_joints.py:90-137``.

``fit_sequence(model, ...)`` reproduces the reference's loop (new Adam per stage,
stage-0 freezing and torso-only visibility, windowed accumulation, temporal term)
on a ``SMALFitter``; ``fused=True`` runs every epoch as one ``FusedFit.step``.
``image_exporter`` (data_io.ResultExporter) turns on the reference's visualisation hook: a collage, the
parameter pickle and the mesh of every frame each ``vis_frequency`` epochs and once more at the end, under
the names the reference uses (optimize_to_joints.py:113-115,139-144).
"""
from __future__ import annotations

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
            model._sync_visibility(0, n)
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
