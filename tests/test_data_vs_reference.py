"""CPU, build container only: the ingest helpers (row 8f-1) against the UNMODIFIED reference loaders imported in place --
smal_fitter/data_loader.py (load_stanford_sequence, load_badja_sequence) and smal_fitter/utils.py (crop_to_silhouette).
Substituted, because absent from the image: imageio.imread (cv2), pycocotools' RLE decoder (the product's own decoder,
which tests/test_data_io.py checks against an independent encoder), nibabel (unused on this path), and the removed
`np.float` alias.  Runs in a child process; skipped where the checkout does not exist.
"""
import json
import os
import subprocess
import sys

import pytest

REF = os.environ.get("SMALIFY_REF", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys, tempfile, types
import numpy as np
import torch
import cv2
repo, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
from smalify_b200 import data_io

def stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod

stub("imageio", imread=lambda p: cv2.imread(p)[:, :, ::-1])
stub("pycocotools")
stub("pycocotools.mask", decode=lambda rle: data_io.decode_coco_rle(rle["counts"], rle["size"][0], rle["size"][1]))
stub("nibabel", eulerangles=types.ModuleType("eulerangles"))
if not hasattr(np, "float"):
    np.float = float                       # data_loader.py:63 uses the alias numpy removed
os.chdir(os.path.join(ref, "smal_fitter"))
sys.path.insert(0, ref); sys.path.insert(0, os.path.join(ref, "smal_fitter"))
import warnings
warnings.simplefilter("ignore")
import data_loader as ref_loader

def diff(a, b):
    (ra, sa, ja, va), na = a
    (rb, sb, jb, vb), nb = b
    assert na == nb, (na, nb)
    assert ra.shape == rb.shape and sa.shape == sb.shape and ja.shape == jb.shape and va.shape == vb.shape, (ra.shape, rb.shape, ja.shape, jb.shape, va.shape, vb.shape)
    return {"rgb": float((ra - rb).abs().max()), "sil": float((sa - sb).abs().max()), "joints": float((ja - jb).abs().max()),
            "vis": float((va.float() - vb.float()).abs().max())}

out = {"stanford": {}, "badja": {}}
sdir = os.path.join(ref, "data", "StanfordExtra")
names = [e["img_path"] for e in json.load(open(os.path.join(sdir, "StanfordExtra_sample.json")))]
for name in names:
    for crop in (96, 256) if name.endswith("n02099601_176.jpg") else (96,):
        out["stanford"]["%s@%d" % (name, crop)] = diff(ref_loader.load_stanford_sequence(sdir, name, crop), data_io.load_stanford_sequence(sdir, name, crop))

# a synthetic BADJA directory: 4 annotated frames, one of them without its segmentation file (skipped by both loaders)
rng = np.random.default_rng(0)
bdir = tempfile.mkdtemp()
os.makedirs(os.path.join(bdir, "joint_annotations")); os.makedirs(os.path.join(bdir, "v", "rgb")); os.makedirs(os.path.join(bdir, "v", "seg"))
ann = []
for i in range(4):
    h, w = 90 + 10 * i, 140
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    seg = np.zeros((h, w, 3), np.uint8)
    cv2.ellipse(seg, (60 + 5 * i, 45), (30 + i, 18), 10 * i, 0, 360, (255, 255, 255), -1)
    cv2.imwrite(os.path.join(bdir, "v", "rgb", "%04d.png" % i), rgb)
    if i != 2:
        cv2.imwrite(os.path.join(bdir, "v", "seg", "%04d.png" % i), seg)
    ann.append({"image_path": "v/rgb/%04d.png" % i, "segmentation_path": "v/seg/%04d.png" % i,
                "joints": rng.integers(0, 90, size=(37, 2)).tolist(), "visibility": (rng.random(37) > 0.4).tolist()})
json.dump(ann, open(os.path.join(bdir, "joint_annotations", "toy.json"), "w"))
for label, rng_ in (("all", None), ("range", range(0, 2))):
    out["badja"][label] = diff(ref_loader.load_badja_sequence(bdir, "toy", 64, image_range=rng_), data_io.load_badja_sequence(bdir, "toy", 64, image_range=rng_))
    out["badja"][label]["frames"] = int(data_io.load_badja_sequence(bdir, "toy", 64, image_range=rng_)[0][0].shape[0])
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smal_fitter")), reason="needs the SMALify checkout (build container only)")
def test_loaders_equal_the_reference_loaders():
    res = subprocess.run([sys.executable, "-c", CHILD, REPO, REF], capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert len(out["stanford"]) == 26                     # the 25 sample images of the checkout, config 1's image at two crop sizes
    for name, d in out["stanford"].items():
        assert d == {"rgb": 0.0, "sil": 0.0, "joints": 0.0, "vis": 0.0}, (name, d)
    assert out["badja"]["all"]["frames"] == 3 and out["badja"]["range"]["frames"] == 2
    for label, d in out["badja"].items():
        assert (d["rgb"], d["sil"], d["joints"], d["vis"]) == (0.0, 0.0, 0.0, 0.0), (label, d)
