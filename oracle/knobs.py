"""Named knobs for every DiffDRR-0.6.0 behaviour restated from memory.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Flip a knob here AND the mirrored constant
in xvr_b200/_conventions.py when a genuine diffdrr install shows a different behaviour; the
parity suite then re-checks kernels against the corrected oracle.
"""

# A1: A.compose(B) applies A first, then B  ->  B.matrix @ A.matrix
COMPOSE_APPLIES_SELF_FIRST = True

# A2: convert(rot, xyz) places the camera centre in the ROTATED frame: matrix = [R | R @ xyz], so that changing
# the angles orbits the C-arm about the isocenter (xvr's sampler draws ty in [700, 900] mm with +-45 deg
# rotations and expects the volume to stay in view; /root/reference/src/xvr/utils/ants.py:71-82 rebuilds a pose
# as make_matrix(R, R @ (A^-1 t)), i.e. make_matrix itself is a plain [R | t] assembler).
# RigidTransform.convert returns xyz = R^T t.  (se3_log_map carries its own translation coupling.)
CONVERT_TRANSLATION_IN_ROTATED_FRAME = True

# A3: detector pixel (row i, col j) -> camera-frame point
#   x = DET_SIGN_S * (j - W//2 + off_w) * delx + x0     (sign flipped by reverse_x_axis)
#   y = DET_SIGN_T * (i - H//2 + off_h) * dely + y0
#   z = sdd
DET_SIGN_S = 1.0
DET_SIGN_T = 1.0

# A6: trilinear renderer
TRILINEAR_N_POINTS = 500
TRILINEAR_ALIGN_CORNERS = True
# Riemann step multiplying the sample sum: "span/(n-1)", "span/n" with span = alphamax-alphamin,
# or "1/n" (whole segment)
TRILINEAR_STEP = "span/(n-1)"
RENDER_EPS = 1e-8

# A7: Siddon renderer.  Planes of axis a sit at i - VOXEL_SHIFT for i in [0, shape_a];
# the midpoint is normalised with 2*(x + VOXEL_SHIFT)/shape - 1 and looked up with
# grid_sample(mode="nearest", align_corners=False)
SIDDON_VOXEL_SHIFT_DEFAULT = 0.5
SIDDON_ALIGN_CORNERS = False

# A9: HU -> density thresholds
HU_AIR = -800.0
HU_BONE = 350.0

# A10/A11: NCC
NCC_EPS = 1e-5

# A12
GEODESIC_EPS = 1e-6
