"""Host-side data ingest and result export around the fitting path.

Restates, without imageio / pycocotools / trimesh (none of them is available offline):
  * ``decode_coco_rle``        pycocotools ``mask.decode`` for the compressed-string RLE the StanfordExtra
                               JSON stores (smal_fitter/data_loader.py:86-96)
  * ``crop_to_silhouette``     smal_fitter/utils.py:5-36
  * ``load_stanford_entry``    smal_fitter/data_loader.py:71-127 (one image -> the 4-tuple SMALFitter takes)
  * ``load_badja_sequence``    smal_fitter/data_loader.py:21-69
  * ``ResultExporter``         the per-frame ``st{stage}_ep{epoch}.pkl`` / ``.ply`` files of
                               ImageExporter.export (smal_fitter/optimize_to_joints.py:25-53) that
                               ``SMALFitter.load_checkpoint`` and generate_video.py read back
One-time CPU work, deliberately plain numpy / cv2: it is not part of the timed hot path.
"""
from __future__ import annotations

import json
import os
import pickle as pkl

import numpy as np
import torch

from . import constants as K


# ------------------------------------------------------------------------------------------------
def decode_coco_rle(counts: str | bytes, height: int, width: int) -> np.ndarray:
    """COCO compressed RLE -> (H, W) uint8 mask.  The string is a LEB128-like stream (6 bits per
    character, offset 48, bit 5 = continuation, bit 4 of the last chunk = sign) of run lengths, every
    run after the third stored as a difference to the run two places back; runs alternate 0/1 in
    column-major order."""
    s = counts.encode("ascii") if isinstance(counts, str) else bytes(counts)
    runs = []
    p = 0
    while p < len(s):
        x = 0
        k = 0
        more = True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(runs) > 2:
            x += runs[-2]
        runs.append(x)
    total = int(np.sum(runs))
    if total != height * width or min(runs, default=0) < 0:
        raise ValueError(f"RLE covers {total} pixels, image has {height * width}")
    flat = np.zeros(height * width, dtype=np.uint8)
    pos = 0
    val = 0
    for r in runs:
        if val:
            flat[pos:pos + r] = 1
        pos += r
        val ^= 1
    return flat.reshape(width, height).T.copy()       # column-major


def crop_to_silhouette(sil_img: np.ndarray, rgb_img: np.ndarray, joints: np.ndarray, target_size: int):
    """Square crop around the silhouette's bounding box (side 1.05 x the larger extent), resized to
    ``target_size``; joints (row, col[, vis]) are mapped into the crop.  Behaviour of utils.py:5-36."""
    import cv2
    assert sil_img.ndim == 2, "Silhouette image is not HxW"
    assert rgb_img.ndim == 3, "RGB image is not HxWx3"
    h, w = sil_img.shape
    pad_sil = np.zeros((h * 4, w * 4))
    pad_rgb = np.ones((h * 4, w * 4, 3))
    pad_sil[h * 2:h * 3, w * 2:w * 3] = sil_img
    pad_rgb[h * 2:h * 3, w * 2:w * 3] = rgb_img
    ys, xs = np.where(pad_sil > 0)
    y_min, y_max, x_min, x_max = ys.min(), ys.max(), xs.min(), xs.max()
    half = int(1.05 * (max(x_max - x_min, y_max - y_min) / 2))
    cy = y_min + int((y_max - y_min) / 2)
    cx = x_min + int((x_max - x_min) / 2)
    sq_sil = pad_sil[cy - half:cy + half, cx - half:cx + half]
    sq_rgb = pad_rgb[cy - half:cy + half, cx - half:cx + half]
    sil_r = cv2.resize(sq_sil, (target_size, target_size), interpolation=cv2.INTER_NEAREST)
    rgb_r = cv2.resize(sq_rgb, (target_size, target_size))
    out = np.zeros_like(joints, dtype=np.float64)
    out[:, 0] = joints[:, 0] + h * 2 - (cy - half)
    out[:, 1] = joints[:, 1] + w * 2 - (cx - half)
    out = out * (target_size / (half * 2.0))
    return sil_r, rgb_r, out


def load_stanford_entry(entry: dict, crop_size: int = K.CROP_SIZE, image: np.ndarray | None = None):
    """One StanfordExtra JSON entry -> ((rgb, sil, joints, visibility), [file name]).  ``image`` is the
    HxWx3 uint8 picture; without it a white canvas stands in (only the visualisation uses the RGB)."""
    h, w = int(entry["img_height"]), int(entry["img_width"])
    seg = decode_coco_rle(entry["seg"], h, w)
    rgb = (np.asarray(image, dtype=np.float64) / 255.0) if image is not None else np.ones((h, w, 3))
    raw = np.concatenate([np.asarray(entry["joints"], dtype=np.float64), [[0.0, 0.0, 0.0]]], axis=0)   # + tail-mid dummy
    sil_img, rgb_img, landmarks = crop_to_silhouette(seg, rgb, raw[:, [1, 0]], crop_size)
    data = (torch.from_numpy(rgb_img).float()[None].permute(0, 3, 1, 2).clamp(0, 1),
            torch.from_numpy(sil_img).float()[None, None],
            torch.from_numpy(landmarks).float()[:, :2][None],
            torch.from_numpy(raw).float()[:, -1][None])
    return data, [os.path.basename(entry["img_path"])]


def load_stanford_sequence(stanford_extra_dir: str, image_name: str, crop_size: int = K.CROP_SIZE):
    import cv2
    with open(os.path.join(stanford_extra_dir, "StanfordExtra_sample.json")) as f:
        entries = {e["img_path"]: e for e in json.load(f)}
    entry = entries[image_name]
    path = os.path.join(stanford_extra_dir, "sample_imgs", entry["img_path"])
    img = cv2.imread(path)
    if img is not None:
        img = img[:, :, ::-1]
    return load_stanford_entry(entry, crop_size, img)


def load_badja_sequence(badja_dir: str, sequence_name: str, crop_size: int = K.CROP_SIZE, image_range=None):
    """data_loader.py:21-69: frames whose rgb or segmentation file is missing are skipped."""
    import cv2
    with open(os.path.join(badja_dir, "joint_annotations", f"{sequence_name}.json")) as f:
        ann = json.load(f)
    if image_range is not None:
        ann = [ann[i] for i in image_range]
    cls = np.asarray(K.BADJA_ANNOTATED_CLASSES)
    rgbs, sils, joints, vis, names = [], [], [], [], []
    for a in ann:
        fn, sn = os.path.join(badja_dir, a["image_path"]), os.path.join(badja_dir, a["segmentation_path"])
        if not (os.path.exists(fn) and os.path.exists(sn)):
            continue
        rgb = cv2.imread(fn)[:, :, ::-1] / 255.0
        seg = cv2.imread(sn)[:, :, ::-1][:, :, 0] / 255.0
        seg = cv2.resize(seg, (rgb.shape[1], rgb.shape[0]), interpolation=cv2.INTER_NEAREST)
        lm = np.asarray(a["joints"], dtype=np.float64)[cls]
        s, r, j = crop_to_silhouette(seg, rgb, lm, crop_size)
        rgbs.append(r); sils.append(s); joints.append(j)
        vis.append(np.asarray(a["visibility"], dtype=bool)[cls] & (cls != -1))
        names.append(os.path.basename(a["image_path"]))
    if not names:
        raise FileNotFoundError(f"no BADJA frame of {sequence_name} has both its rgb and segmentation file under {badja_dir}")
    data = (torch.from_numpy(np.stack(rgbs)).float().permute(0, 3, 1, 2).clamp(0, 1),
            torch.from_numpy(np.stack(sils)).float()[:, None],
            torch.from_numpy(np.stack(joints)).float(),
            torch.from_numpy(np.stack(vis).astype(np.float32)))
    return data, names


# ------------------------------------------------------------------------------------------------
def write_ply(path: str, vertices: np.ndarray, faces: np.ndarray) -> None:
    """Binary little-endian PLY triangle mesh (what trimesh.export writes for a .ply)."""
    v = np.ascontiguousarray(vertices, dtype="<f4")
    f = np.ascontiguousarray(faces, dtype="<i4")
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n").encode("ascii")
    rec = np.empty(len(f), dtype=[("n", "u1"), ("idx", "<i4", (3,))])
    rec["n"] = 3
    rec["idx"] = f
    with open(path, "wb") as fh:
        fh.write(header)
        fh.write(v.tobytes())
        fh.write(rec.tobytes())


class ResultExporter:
    """Directory layout and file names of ImageExporter (optimize_to_joints.py:25-53)."""

    def __init__(self, output_dir: str, filenames):
        os.makedirs(output_dir, exist_ok=True)
        self.output_dirs = []
        for fn in filenames:
            d = os.path.join(output_dir, os.path.splitext(fn)[0])
            os.makedirs(d, exist_ok=True)
            self.output_dirs.append(d)
        self.stage_id = 0
        self.epoch_name = "0"

    def export(self, collage_np, batch_id, global_id, img_parameters, vertices, faces):
        """ImageExporter.export (optimize_to_joints.py:43-53): collage png + parameter pkl + mesh ply of one frame."""
        import cv2
        stem = os.path.join(self.output_dirs[global_id], "st{0}_ep{1}".format(self.stage_id, self.epoch_name))
        cv2.imwrite(stem + ".png", np.ascontiguousarray(collage_np[:, :, ::-1]))        # RGB -> BGR for cv2
        with open(stem + ".pkl", "wb") as f:
            pkl.dump(img_parameters, f)
        v = vertices[batch_id]
        write_ply(stem + ".ply", v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v), np.asarray(faces))

    def export_fitter(self, fitter, write_mesh: bool = True):
        """The pkl (5 keys) and ply of every frame at the fitter's current parameters."""
        verts = fitter.vertices().cpu().numpy() if write_mesh else None
        faces = np.asarray(fitter.constants.faces)
        stem = "st{0}_ep{1}".format(self.stage_id, self.epoch_name)
        for i, d in enumerate(self.output_dirs):
            with open(os.path.join(d, stem + ".pkl"), "wb") as f:
                pkl.dump(fitter.export_parameters(i), f)
            if write_mesh:
                write_ply(os.path.join(d, stem + ".ply"), verts[i], faces)
