"""CPU, build container only: the CALL SITES of the unpinned half.  PyTorch3D 0.2.5 is not available, so its algorithm is
restated (oracle/smal_oracle.py) -- but how the reference configures and calls it is executable: the reference's
smal_fitter/p3d_renderer.py is imported unmodified against a *recording* stand-in of the `pytorch3d` package, and every
argument it passes (camera placement, blend parameters, rasterisation settings, which output channel and which column
order it takes) is compared with the constants the oracle and the product are built on.  Child process; skipped where
the checkout does not exist.
"""
import json
import os
import subprocess
import sys

import pytest

REF = os.environ.get("SMALIFY_REF", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys, types
import numpy as np
import torch
repo, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
from smalify_b200 import constants as K, visualization as V
from oracle import smal_oracle as O

calls = []
def recorder(name):
    class Rec:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k
            calls.append((name, a, k))
            for kk, vv in k.items():
                setattr(self, kk, vv)
        def __call__(self, mesh):                                   # MeshRenderer(...)(mesh)
            calls.append((name + ".__call__", (mesh,), {}))
            B = mesh.kwargs["verts"].shape[0]
            S = self.kwargs["rasterizer"].kwargs["raster_settings"].kwargs["image_size"]
            out = torch.zeros(B, S, S, 4)
            out[..., 3] = 0.25                                       # alpha channel marker
            out[..., 0] = 0.5
            return out
        def transform_points_screen(self, points, screen_size):     # cameras.transform_points_screen
            calls.append((name + ".transform_points_screen", (points, screen_size), {}))
            out = torch.zeros(points.shape[0], points.shape[1], 3)
            out[..., 0] = 1.0                                        # x (column)
            out[..., 1] = 2.0                                        # y (row)
            return out
    Rec.__name__ = name
    return Rec

def look_at_view_transform(*a, **k):
    calls.append(("look_at_view_transform", a, k))
    return "R", "T"

names = ["OpenGLPerspectiveCameras", "RasterizationSettings", "MeshRenderer", "MeshRasterizer", "BlendParams", "PointLights",
         "HardPhongShader", "SoftSilhouetteShader", "Materials", "Textures", "look_at_rotation"]
p3d = types.ModuleType("pytorch3d"); sys.modules["pytorch3d"] = p3d
ren = types.ModuleType("pytorch3d.renderer"); sys.modules["pytorch3d.renderer"] = ren
for n in names:
    setattr(ren, n, recorder(n))
ren.look_at_view_transform = look_at_view_transform
st = types.ModuleType("pytorch3d.structures"); st.Meshes = recorder("Meshes"); sys.modules["pytorch3d.structures"] = st
io_ = types.ModuleType("pytorch3d.io"); io_.load_objs_as_meshes = None; sys.modules["pytorch3d.io"] = io_
ut = types.ModuleType("utils"); ut.perspective_proj_withz = None; sys.modules["utils"] = ut
os.chdir(ref)
sys.path.insert(0, ref); sys.path.insert(0, os.path.join(ref, "smal_fitter"))
import p3d_renderer
r = p3d_renderer.Renderer(256, "cpu")
B, V_, F_ = 2, 5, 3
sil, proj = r(torch.zeros(B, V_, 3), torch.zeros(B, 25, 3), torch.zeros(B, F_, 3, dtype=torch.long))
sil3, proj3, col = r(torch.zeros(B, V_, 3), torch.zeros(B, 25, 3), torch.zeros(B, F_, 3, dtype=torch.long), render_texture=True)

def first(name):
    return next(c for c in calls if c[0] == name)
def all_(name):
    return [c for c in calls if c[0] == name]
lav = first("look_at_view_transform")
cam = first("OpenGLPerspectiveCameras")
blend = first("BlendParams")
rs = all_("RasterizationSettings")
lights = first("PointLights")
tex = first("Textures")
scr = first("OpenGLPerspectiveCameras.transform_points_screen")
out = {
    "look_at_args": [float(x) for x in lav[1]], "look_at_kwargs": sorted(lav[2]),
    "camera_kwargs": sorted(cam[2]),
    "blend": {k: float(v) for k, v in blend[2].items()},
    "soft": {k: (float(v) if not isinstance(v, int) else v) for k, v in rs[0][2].items()},
    "hard": {k: (float(v) if not isinstance(v, int) else v) for k, v in rs[1][2].items()},
    "light_location": lights[2]["location"],
    "texture_rgb": [float(x) for x in tex[2]["verts_rgb"][0, 0]],
    "sil_shape": list(sil.shape), "sil_value": float(sil.max()),         # 0.25 == the LAST channel was taken
    "proj_first_is_row": float(proj[0, 0, 0]) == 2.0 and float(proj[0, 0, 1]) == 1.0,
    "screen_size": [float(x) for x in scr[1][1][0]],
    "color_shape": list(col.shape),
    "ours": {"distance": K.CAMERA_DISTANCE, "oracle_distance": O.CAMERA_DISTANCE, "sigma": K.BLEND_SIGMA, "oracle_sigma": O.SIGMA,
             "blur": K.BLUR_RADIUS, "oracle_blur": O.BLUR_RADIUS, "k": K.FACES_PER_PIXEL, "mesh_color": list(V.MESH_COLOR)},
}
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smal_fitter")), reason="needs the SMALify checkout (build container only)")
def test_renderer_configuration_equals_the_reference_call_sites():
    res = subprocess.run([sys.executable, "-c", CHILD, REPO, REF], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    ours = out["ours"]
    # camera: look_at_view_transform(dist, elev = 0, azim = 0), OpenGLPerspectiveCameras with nothing but R, T (defaults: fov 60)
    assert out["look_at_args"] == [ours["distance"], 0.0, 0.0] == [ours["oracle_distance"], 0.0, 0.0] and out["look_at_kwargs"] == ["device"]
    assert out["camera_kwargs"] == ["R", "T", "device"]
    # blend and soft rasterisation settings (p3d_renderer.py:26-31): sigma, blur radius, K -- and no other setting touched
    assert out["blend"] == {"sigma": ours["sigma"], "gamma": 1e-4} and ours["sigma"] == ours["oracle_sigma"]
    assert sorted(out["soft"]) == ["blur_radius", "faces_per_pixel", "image_size"]
    assert out["soft"]["image_size"] == 256 and out["soft"]["faces_per_pixel"] == ours["k"] == 100
    assert abs(out["soft"]["blur_radius"] - ours["blur"]) <= 1e-18 and abs(out["soft"]["blur_radius"] - ours["oracle_blur"]) <= 1e-18
    # the silhouette is the last channel of the shader's output as (B, 1, S, S); keypoints come back as (row, col) in pixels
    assert out["sil_shape"] == [2, 1, 256, 256] and out["sil_value"] == 0.25
    assert out["proj_first_is_row"] and out["screen_size"] == [256.0, 256.0]
    # colour pass of the visualisation (rows 8f-3): hard rasterisation, one face per pixel, a point light at (0, 0, 3), flat mesh colour
    assert out["hard"] == {"image_size": 256, "blur_radius": 0.0, "faces_per_pixel": 1}
    assert out["light_location"] == [[0.0, 0.0, 3.0]]
    assert all(abs(a - b) < 1e-7 for a, b in zip(out["texture_rgb"], ours["mesh_color"]))
    assert out["color_shape"] == [2, 3, 256, 256]
