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


CHILD_CKPT = r'''
import json, os, pickle, sys, tempfile, types
import numpy as np
import torch
repo, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo); sys.path.insert(0, os.path.join(repo, "tests"))
from smalify_b200 import constants as K, data_io, model_io, synthetic
from smalify_b200.model_io import _ChStub

def stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod

saved = {}
stub("chumpy", Ch=_ChStub); stub("chumpy.ch", Ch=_ChStub)
stub("matplotlib"); stub("matplotlib.pyplot"); stub("trimesh", Trimesh=lambda vertices, faces, process: types.SimpleNamespace(export=lambda p: saved.setdefault("ply", []).append((p, np.asarray(vertices), np.asarray(faces)))))
stub("imageio", imsave=lambda p, a: saved.setdefault("png", []).append((p, a.shape)))
stub("data_loader", load_badja_sequence=None, load_stanford_sequence=None)
stub("draw_smal_joints", SMALJointDrawer=type("SMALJointDrawer", (), {}))
stub("utils", eul_to_axis=lambda e: np.asarray(K.GLOBAL_ROT_INIT, dtype=np.float64))
stub("p3d_renderer", Renderer=type("Renderer", (torch.nn.Module,), {"__init__": lambda self, s, d: torch.nn.Module.__init__(self)}))
os.chdir(ref)
sys.path.insert(0, ref); sys.path.insert(0, os.path.join(ref, "smal_fitter"))
import warnings
warnings.simplefilter("ignore")
from smal_fitter import SMALFitter
import optimize_to_joints as ref_loop

N, S = 3, 16
g = torch.Generator().manual_seed(3)
data = (torch.rand(N, 3, S, S, generator=g), torch.zeros(N, 1, S, S), torch.zeros(N, 25, 2), torch.ones(N, 25))
names = ["%04d.png" % i for i in range(N)]
want = {"global_rotation": torch.randn(N, 3, generator=g), "joint_rotations": torch.randn(N, 34, 3, generator=g),
        "betas": torch.randn(20, generator=g), "log_betascale": torch.randn(6, generator=g), "trans": torch.randn(N, 3, generator=g)}

class Fake:                     # what SMALFitter.export_parameters / vertices of the product return (tests/test_gpu_io.py runs the real one)
    constants = types.SimpleNamespace(faces=np.array([[0, 1, 2], [2, 1, 3]]))
    def vertices(self):
        return torch.arange(N * 4 * 3, dtype=torch.float32).reshape(N, 4, 3)
    def export_parameters(self, i):
        return {"global_rotation": want["global_rotation"][i].numpy(), "joint_rotations": want["joint_rotations"][i].numpy(),
                "betas": want["betas"].numpy(), "log_betascale": want["log_betascale"].numpy(), "trans": want["trans"][i].numpy()}

out = {}
# 1. the product's exporter writes, the reference's load_checkpoint reads (smal_fitter.py:192-207)
d1 = tempfile.mkdtemp()
ex = data_io.ResultExporter(d1, names)
ex.stage_id, ex.epoch_name = 10, "0"
ex.export_fitter(Fake())
model = SMALFitter("cpu", data, N, 1, True)
with torch.no_grad():                      # (its in-place row writes on leaf parameters need this under current torch)
    model.load_checkpoint(d1, "st10_ep0")
out["ref_reads_ours"] = {"global_rotation": float((model.global_rotation.detach() - want["global_rotation"]).abs().max()),
                         "joint_rotations": float((model.joint_rotations.detach() - want["joint_rotations"]).abs().max()),
                         "trans": float((model.trans.detach() - want["trans"]).abs().max()),
                         "betas": float((model.betas.detach() - want["betas"]).abs().max()),
                         "log_beta_scales": float((model.log_beta_scales.detach() - want["log_betascale"]).abs().max())}
# 2. the reference's ImageExporter writes (optimize_to_joints.py:25-53), the product's files are compared with it
d2 = tempfile.mkdtemp()
rex = ref_loop.ImageExporter(d2, names)
rex.stage_id, rex.epoch_name = 10, "0"
verts = Fake().vertices()
for i in range(N):
    rex.export(np.zeros((S, S * 5, 3), np.uint8), i, i, Fake().export_parameters(i), verts, Fake.constants.faces)
    ex.export(np.zeros((S, S * 5, 3), np.uint8), i, i, Fake().export_parameters(i), verts, Fake.constants.faces)
same_tree = sorted(os.path.relpath(os.path.join(r, f), d2) for r, _, fs in os.walk(d2) for f in fs)
ours_tree = sorted(os.path.relpath(os.path.join(r, f), d1) for r, _, fs in os.walk(d1) for f in fs)
out["ref_tree"] = same_tree
out["ours_tree"] = ours_tree
pk = []
for i in range(N):
    a = pickle.load(open(os.path.join(d2, "%04d" % i, "st10_ep0.pkl"), "rb"))
    b = pickle.load(open(os.path.join(d1, "%04d" % i, "st10_ep0.pkl"), "rb"))
    pk.append(sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype and a[k].shape == b[k].shape for k in a))
out["pkl_equal"] = pk
out["ref_png_ply_calls"] = [len(saved.get("png", [])), len(saved.get("ply", []))]
out["ref_ply_paths"] = [os.path.relpath(p, d2) for p, _, _ in saved.get("ply", [])]
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smal_fitter")), reason="needs the SMALify checkout (build container only)")
def test_checkpoint_wire_format_against_the_reference():
    """Row 8f-2: the reference's `SMALFitter.load_checkpoint` reads what the product's exporter wrote, and the product's
    exporter produces the file tree and the parameter pickles of the reference's `ImageExporter.export`."""
    res = subprocess.run([sys.executable, "-c", CHILD_CKPT, REPO, REF], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    r = out["ref_reads_ours"]
    assert r["global_rotation"] == 0.0 and r["joint_rotations"] == 0.0 and r["trans"] == 0.0, r
    assert r["betas"] <= 1e-6 and r["log_beta_scales"] <= 1e-6, r          # (the reference averages the frames' copies: float32 mean)
    assert out["pkl_equal"] == [True, True, True]
    # the reference wrote its pkl files itself; png / ply went through the substituted imageio / trimesh: same names as ours
    expect = sorted("%04d/st10_ep0.%s" % (i, e) for i in range(3) for e in ("pkl", "png", "ply"))
    assert out["ours_tree"] == expect
    assert out["ref_tree"] == sorted(p for p in expect if p.endswith(".pkl"))
    assert out["ref_png_ply_calls"] == [3, 3] and sorted(out["ref_ply_paths"]) == sorted(p for p in expect if p.endswith(".ply"))


CHILD_MARKERS = r'''
import json, os, sys, types
import numpy as np
import torch
repo, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
from smalify_b200 import constants as K, visualization as V
for name in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.utils"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["torchvision.utils"].make_grid = None
os.chdir(ref)
sys.path.insert(0, ref); sys.path.insert(0, os.path.join(ref, "smal_fitter"))
import config
from draw_smal_joints import SMALJointDrawer
g = torch.Generator().manual_seed(5)
img = torch.rand(3, 3, 96, 96, generator=g)
lm = torch.randint(0, 96, (3, 25, 2), generator=g)            # integer (row, col): current cv2 rejects the float32 the reference passes
vis = torch.rand(3, 25, generator=g) > 0.3
out = {"tables": bool(np.array_equal(np.array(config.MARKER_COLORS), np.array(V.MARKER_COLORS)) and list(config.MARKER_TYPE) == list(V.MARKER_TYPE)),
       "mesh_color": bool(np.allclose(np.array(config.MESH_COLOR) / 255.0, np.array(V.MESH_COLOR))),
       "canonical": list(config.CANONICAL_MODEL_JOINTS) == list(K.CANONICAL_MODEL_JOINTS), "torso": list(config.TORSO_JOINTS) == list(K.TORSO_JOINTS),
       "badja_classes": list(config.BADJA_ANNOTATED_CLASSES) == list(K.BADJA_ANNOTATED_CLASSES),
       "n_pose_betas": [config.N_POSE == K.N_POSE, config.N_BETAS == K.N_BETAS], "crop_window_vis": [config.CROP_SIZE == K.CROP_SIZE, config.WINDOW_SIZE == K.WINDOW_SIZE, config.VIS_FREQUENCY == K.VIS_FREQUENCY]}
for label, v in (("with_visibility", vis), ("all_visible", None)):
    a = SMALJointDrawer.draw_joints(img, lm, visible=v)
    b = V.draw_joints(img, lm.float(), visible=v)
    out[label] = float((a - b).abs().max())
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smal_fitter")), reason="needs the SMALify checkout (build container only)")
def test_joint_markers_and_config_tables_equal_the_reference():
    """Row 8f-3 (host side) and the constant tables of config.py: marker drawing pixel for pixel, marker / colour tables,
    joint index lists, sizes."""
    res = subprocess.run([sys.executable, "-c", CHILD_MARKERS, REPO, REF], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["tables"] and out["mesh_color"] and out["canonical"] and out["torso"] and out["badja_classes"], out
    assert out["n_pose_betas"] == [True, True] and out["crop_window_vis"] == [True, True, True], out
    assert out["with_visibility"] == 0.0 and out["all_visible"] == 0.0, out
