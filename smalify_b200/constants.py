"""Numeric constants of the SMALify fitting path.

These are the *data* the hot path consumes; each block cites where the
reference defines the same values (paths relative to the SMALify checkout).
Nothing here is executable logic of the reference.
"""
from __future__ import annotations

import math

# Mesh / skeleton sizes (SMAL model pickle, config.py:131-132).
N_VERTS = 3889
N_FACES = 7774
N_JOINTS = 35          # global + 34 body joints
N_POSE = 34            # joint_rotations rows (config.py:131)
N_BETAS = 20           # config.py:132
N_LOGSCALE = 6         # smal_fitter.py:61
N_MODEL_JOINTS = 41    # 35 regressed + 6 picked vertices (smal_torch.py:171-184)
N_KEYPOINTS = 25       # len(CANONICAL_MODEL_JOINTS), config.py:77-88

# Vertices appended as keypoints 35..40: nose, chin, r-ear tip, l-ear tip,
# l-eye, r-eye (smal_torch.py:176-183).
PICKED_VERTS = (1863, 26, 2124, 150, 3055, 1097)

# Model-joint index of each of the 25 annotated keypoints (config.py:77-88).
CANONICAL_MODEL_JOINTS = (
    10, 9, 8,
    20, 19, 18,
    14, 13, 12,
    24, 23, 22,
    25, 31,
    33, 34,
    35, 36,
    38, 37,
    39, 40,
    15, 15,
    28,
)

# Keypoints that stay visible during stage 0 (config.py:75).
TORSO_JOINTS = (2, 5, 8, 11, 12, 23)

# BADJA annotation index per keypoint, -1 = not annotated (config.py:91-102).
BADJA_ANNOTATED_CLASSES = (
    14, 13, 12,
    24, 23, 22,
    10, 9, 8,
    20, 19, 18,
    25, 31,
    -1, -1,
    33, -1,
    36, 35,
    -1, -1,
    -1, 15,
    28,
)

# Limb-scale groups (batch_lbs.py:107-121): joint ranges and, per xyz axis,
# which of the 6 log-scale entries multiplies that axis (-1 = none).
SCALE_GROUPS = (
    # (first joint, last joint inclusive, (idx for x, idx for y, idx for z))
    (7, 10, (1, 1, 0)),
    (11, 14, (1, 1, 0)),
    (17, 20, (1, 1, 0)),
    (21, 24, (1, 1, 0)),
    (25, 31, (2, 3, 3)),
    (33, 34, (-1, 4, 5)),
)

# Vertices the reference pins to the symmetry plane (smal_basics.py:9), as
# inclusive runs.
SYMMETRY_AXIS_RUNS = (
    (0, 32), (37, 37), (55, 55), (119, 120), (163, 163), (209, 211),
    (213, 213), (216, 216), (227, 227), (326, 326), (395, 395), (452, 452),
    (578, 578), (910, 910), (959, 959), (964, 964), (975, 977), (1172, 1172),
    (1175, 1176), (1178, 1178), (1194, 1194), (1243, 1243), (1739, 1739),
    (1796, 1840), (1842, 1863), (1870, 1870), (1919, 1919), (1960, 1961),
    (1965, 1965), (1967, 1967), (2003, 2003),
)

# Camera (p3d_renderer.py:22-23: look_at_view_transform(2.7, 0, 0) and the
# default 60 degree OpenGL perspective camera of PyTorch3D 0.2.5).
CAMERA_DISTANCE = 2.7
VIS_FREQUENCY = 100                     # config.py: collage / checkpoint export every this many epochs
CAMERA_FOCAL = 1.7320508075688772       # 1 / tan(fov / 2), OpenGLPerspectiveCameras default fov = 60 degrees
NDC_FOCAL = 1.0 / math.tan(math.radians(60.0) / 2.0)   # sqrt(3)

# Soft silhouette (p3d_renderer.py:26-31).
BLEND_SIGMA = 1e-4
BLUR_RADIUS = math.log(1.0 / 1e-4 - 1.0) * BLEND_SIGMA  # 9.21024e-4 (NDC^2)
FACES_PER_PIXEL = 100
RASTER_EPS = 1e-8       # PyTorch3D kEpsilon (csrc/utils/geometry_utils.cuh)

# Stage schedule, one column of OPT_WEIGHTS per stage (config.py:63-72):
# (w_j2d, w_sil, w_betas, w_pose, w_limit, w_splay, w_temp, iters, lr)
STAGE_SCHEDULE = (
    (25.0, 0.0, 0.0, 0.0, 0.0, 0.0, 500.0, 150, 5e-3),
    (10.0, 500.0, 1.0, 1.0, 100.0, 0.1, 100.0, 400, 5e-3),
    (7.5, 5000.0, 1.0, 1.0, 100.0, 0.1, 100.0, 600, 5e-4),
    (5.0, 5000.0, 1.0, 1.0, 100.0, 0.1, 100.0, 800, 1e-4),
)

ADAM_BETAS = (0.5, 0.999)   # optimize_to_joints.py:96
ADAM_EPS = 1e-8

# Global-rotation initialisation: eul_to_axis([-pi/2, 0, -pi/2])
# (smal_fitter.py:81, utils.py:61-63 via nibabel euler2angle_axis(z,y,x)):
# R = Rx(-pi/2) Ry(0) Rz(-pi/2) = [[0,1,0],[0,0,1],[1,0,0]], a rotation of
# -2pi/3 about (1,1,1)/sqrt(3)  ->  rot-vec = -(2pi/3)/sqrt(3) * (1,1,1).
GLOBAL_ROT_INIT = tuple([-(2.0 * math.pi / 3.0) / math.sqrt(3.0)] * 3)

CROP_SIZE = 256     # config.py:12
WINDOW_SIZE = 10    # config.py:25


# Joint-rotation limits (radians) of smal_fitter/priors/joint_limits_prior.py (`Ranges`, in the order of
# `LimitPrior.parts`: part id p is joint_rotations row p, i.e. skeleton joint p + 1), as
# (x_min, x_max, y_min, y_max, z_min, z_max) per joint.  The reference's table has 32 parts; the 35-joint
# model has 34 pose joints: the two ear joints (rows 32, 33) are left unbounded.  The term is
# commented out in the reference (smal_fitter.py:146-151) and off by default here.
_JOINT_LIMIT_ROWS = (
    (-0.3, 0.3, -1.2, 0.5, -0.1, 0.1),     # pelvis0
    (-0.4, 0.4, -1.0, 0.9, -0.8, 0.8),     # spine
    (-0.4, 0.4, -1.0, 0.9, -0.8, 0.8),     # spine0
    (-0.4, 0.4, -0.5, 1.2, -0.4, 0.4),     # spine1
    (-0.5, 0.5, -0.4, 1.4, -0.5, 0.5),     # spine2
    (-0.5, 0.5, -0.6, 1.4, -0.8, 0.8),     # spine3
    (-0.05, 0.05, -1.3, 0.8, -0.6, 0.6),   # LLeg1
    (-0.05, 0.05, -1.0, 1.1, -0.6, 0.6),   # LLeg2
    (-0.4, 0.1, -0.3, 1.4, -0.7, 0.4),     # LLeg3
    (-0.3, 0.1, -0.4, 1.5, -0.7, 0.3),     # LFoot
    (-0.05, 0.05, -1.3, 0.8, -0.6, 0.6),   # RLeg1
    (-0.05, 0.05, -1.0, 0.9, -0.6, 0.6),   # RLeg2
    (-0.1, 0.4, -0.3, 1.4, -0.4, 0.7),     # RLeg3
    (-0.1, 0.3, -0.4, 1.5, -0.3, 0.7),     # RFoot
    (-0.8, 0.8, -1.0, 1.0, -1.1, 1.1),     # Neck
    (-0.5, 0.5, -1.0, 0.9, -0.9, 0.9),     # Head
    (-0.2, 0.3, -0.5, 0.8, -0.5, 0.4),     # LLegBack1
    (-0.2, 0.3, -0.6, 0.8, -0.6, 0.5),     # LLegBack2
    (-0.3, 0.2, -0.8, 0.2, -0.5, 0.4),     # LLegBack3
    (-0.3, 0.2, -0.3, 1.1, -0.5, 0.3),     # LFootBack
    (-0.3, 0.2, -0.5, 0.8, -0.4, 0.5),     # RLegBack1
    (-0.3, 0.2, -0.6, 0.8, -0.5, 0.6),     # RLegBack2
    (-0.2, 0.3, -0.8, 0.2, -0.4, 0.5),     # RLegBack3
    (-0.2, 0.3, -0.3, 1.1, -0.3, 0.5),     # RFootBack
    (-0.1, 0.1, -1.5, 1.4, -1.2, 1.2),     # Tail1
    (-0.1, 0.1, -1.0, 1.0, -0.8, 0.8),     # Tail2
    (-0.1, 0.1, -1.0, 1.0, -0.8, 0.8),     # Tail3
    (-0.1, 0.1, -1.0, 1.0, -0.8, 0.8),     # Tail4
    (-0.1, 0.1, -1.0, 1.0, -0.8, 0.8),     # Tail5
    (-0.1, 0.1, -1.4, 1.4, -1.0, 1.0),     # Tail6
    (-0.1, 0.1, -0.7, 1.1, -0.9, 0.8),     # Tail7
    (-0.1, 0.1, -1.1, 0.5, -0.1, 0.1),     # Mouth
)


def joint_limits():
    """(min, max) float32 arrays of shape (N_POSE, 3); ears unbounded."""
    import numpy as np
    lo = np.full((N_POSE, 3), -np.inf, np.float32)
    hi = np.full((N_POSE, 3), np.inf, np.float32)
    rows = np.asarray(_JOINT_LIMIT_ROWS, np.float32).reshape(-1, 3, 2)
    lo[: rows.shape[0]] = rows[:, :, 0]
    hi[: rows.shape[0]] = rows[:, :, 1]
    return lo, hi
