"""Conventions of DiffDRR 0.6.0 that the reference tree cannot pin.

xvr delegates its whole hot path to the un-vendored ``diffdrr==0.6.0`` (/root/reference/pyproject.toml:14).
Every behaviour that had to be restated without the package at hand is a named constant here; the test
oracle keeps a mirrored copy (oracle/knobs.py) and tests/test_conventions.py asserts the two agree, so a
correction is made in exactly two places and re-checked by the parity suite.
"""

# RigidTransform.compose: A.compose(B) applies A first, then B -> B.matrix @ A.matrix
COMPOSE_APPLIES_SELF_FIRST = True

# convert(rot, xyz): the camera centre is given in the rotated frame, matrix = [R | R @ xyz] (angles orbit the
# C-arm about the isocenter); RigidTransform.convert returns xyz = R^T t.  make_matrix is a plain assembler
# (/root/reference/src/xvr/utils/ants.py:71-82).  se3_log_map keeps its own translation coupling.
CONVERT_TRANSLATION_IN_ROTATED_FRAME = True

# Detector pixel (row i, col j) -> camera-frame point (x, y, z) =
#   (DET_SIGN_S * (j - W//2 + off_w) * delx + x0, DET_SIGN_T * (i - H//2 + off_h) * dely + y0, sdd)
# with DET_SIGN_S negated by reverse_x_axis.
DET_SIGN_S = 1.0
DET_SIGN_T = 1.0

# Trilinear renderer
TRILINEAR_N_POINTS = 500
TRILINEAR_STEP = "span/(n-1)"  # also "span/n", "1/n"
STEP_MODES = {"span/(n-1)": 0, "span/n": 1, "1/n": 2}
RENDER_EPS = 1e-8

# Siddon renderer: planes of axis a at i - voxel_shift, i in [0, shape_a]
SIDDON_VOXEL_SHIFT_DEFAULT = 0.5

# transform_hu_to_density thresholds
HU_AIR = -800.0
HU_BONE = 350.0

NCC_EPS = 1e-5
GEODESIC_EPS = 1e-6
