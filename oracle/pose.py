"""Oracle restatement of diffdrr.pose (SE(3) parameterisations) -- TEST INFRASTRUCTURE ONLY.

Follows diffdrr 0.6.0 ``diffdrr/pose.py`` (convert, RigidTransform.convert, make_matrix and the
pytorch3d-ported rotation conversions) as pinned by the xvr call sites
/root/reference/src/xvr/model/sampler.py:29-31, model/network.py:49-54, model/trainer.py:335-337,
registrar/base.py:168,201.  Functional style on raw (B,4,4) matrices; PARITY UNPINNED (see
oracle/__init__.py).
"""

import math

import torch

from . import knobs

__all__ = [
    "N_ANGULAR_COMPONENTS",
    "pose_from_params",
    "params_from_pose",
    "make_matrix",
    "compose",
    "invert",
    "apply",
    "euler_angles_to_matrix",
    "matrix_to_euler_angles",
    "quaternion_to_matrix",
    "matrix_to_quaternion",
    "axis_angle_to_matrix",
    "matrix_to_axis_angle",
    "so3_exp_map",
    "so3_log_map",
]

# diffdrr/registration.py N_ANGULAR_COMPONENTS (used at /root/reference/src/xvr/model/network.py:28)
N_ANGULAR_COMPONENTS = {
    "axis_angle": 3,
    "euler_angles": 3,
    "se3_log_map": 3,
    "so3_log_map": 3,
    "quaternion": 4,
    "rotation_6d": 6,
    "rotation_10d": 10,
    "quaternion_adjugate": 10,
}


# ----------------------------------------------------------------------------- elementary
def _axis_rotation(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == "Y":
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    elif axis == "Z":
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    else:
        raise ValueError(axis)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(angles, convention):
    """R = R_c0(a0) @ R_c1(a1) @ R_c2(a2) (pytorch3d convention)."""
    mats = [_axis_rotation(c, a) for c, a in zip(convention, torch.unbind(angles, -1))]
    return mats[0] @ mats[1] @ mats[2]


def _angle_from_tan(axis, other_axis, data, horizontal, tait_bryan):
    i1, i2 = {"X": (2, 1), "Y": (0, 2), "Z": (1, 0)}[axis]
    if horizontal:
        i2, i1 = i1, i2
    even = (axis + other_axis) in ["XY", "YZ", "ZX"]
    if horizontal == even:
        return torch.atan2(data[..., i1], data[..., i2])
    if tait_bryan:
        return torch.atan2(-data[..., i2], data[..., i1])
    return torch.atan2(data[..., i2], -data[..., i1])


def matrix_to_euler_angles(matrix, convention):
    idx = {"X": 0, "Y": 1, "Z": 2}
    i0, i2 = idx[convention[0]], idx[convention[2]]
    tait_bryan = i0 != i2
    if tait_bryan:
        sign = -1.0 if (i0 - i2) in (-1, 2) else 1.0
        central = torch.asin(matrix[..., i0, i2] * sign)
    else:
        central = torch.acos(matrix[..., i0, i0])
    o = (
        _angle_from_tan(convention[0], convention[1], matrix[..., i2], False, tait_bryan),
        central,
        _angle_from_tan(convention[2], convention[1], matrix[..., i0, :], True, tait_bryan),
    )
    return torch.stack(o, -1)


def quaternion_to_matrix(q):
    """(w,x,y,z) real-first, not required to be normalised (scaled by 2/|q|^2)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(matrix):
    batch = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch + (9,)), -1)
    q_abs = _sqrt_positive_part(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
            -1,
        )
    )
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    )
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = q_abs.argmax(-1)
    out = torch.gather(cand, -2, best[..., None, None].expand(batch + (1, 4)))[..., 0, :]
    # standardise to non-negative real part
    return torch.where(out[..., 0:1] < 0, -out, out)


def _hat(v):
    x, y, z = torch.unbind(v, -1)
    zero = torch.zeros_like(x)
    return torch.stack((zero, -z, y, z, zero, -x, -y, x, zero), -1).reshape(v.shape[:-1] + (3, 3))


def so3_exp_map(log_rot, eps=1e-4):
    """Rodrigues: R = I + sin(t)/t K + (1-cos t)/t^2 K^2 with t clamped like pytorch3d."""
    nrms = (log_rot * log_rot).sum(-1)
    theta = nrms.clamp_min(eps).sqrt()
    fac1 = (theta.sin() / theta)[..., None, None]
    fac2 = ((1 - theta.cos()) / (theta * theta))[..., None, None]
    K = _hat(log_rot)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)
    return eye + fac1 * K + fac2 * (K @ K)


def so3_log_map(R, eps=1e-4):
    """Inverse of so3_exp_map (pytorch3d so3_log_map: acos of the clamped trace)."""
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    cos = ((tr - 1.0) * 0.5).clamp(-1.0 + 1e-7, 1.0 - 1e-7)
    phi = torch.acos(cos)
    sin = phi.sin()
    fac = torch.where(sin.abs() > 0.5 * eps, phi / (2.0 * sin.clamp_min(1e-12)), 0.5 + phi * phi / 12.0)
    skew = fac[..., None, None] * (R - R.transpose(-1, -2))
    return torch.stack((skew[..., 2, 1], skew[..., 0, 2], skew[..., 1, 0]), -1)


def axis_angle_to_matrix(aa):
    return so3_exp_map(aa, eps=1e-12)


def matrix_to_axis_angle(R):
    q = matrix_to_quaternion(R)
    norms = q[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    angles = 2 * half
    small = angles.abs() < 1e-6
    s = torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    return q[..., 1:] / s


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = torch.nn.functional.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), -2)


def matrix_to_rotation_6d(R):
    return R[..., :2, :].clone().reshape(R.shape[:-2] + (6,))


def _sym4(vec):
    idx, jdx = torch.triu_indices(4, 4)
    A = torch.zeros(vec.shape[:-1] + (4, 4), dtype=vec.dtype, device=vec.device)
    A[..., idx, jdx] = vec
    A[..., jdx, idx] = vec
    return A


def quaternion_adjugate_to_quaternion(vec):
    """10-vector -> symmetric 4x4 (the adjugate q q^T up to scale) -> its max-norm column, normalised."""
    A = _sym4(vec)
    col_norm = A.norm(dim=-2)  # (B,4)
    best = col_norm.argmax(-1)
    col = torch.gather(A, -1, best[..., None, None].expand(A.shape[:-1] + (1,)))[..., 0]
    return col / col_norm.gather(-1, best[..., None])


def quaternion_to_quaternion_adjugate(q):
    A = q[..., :, None] * q[..., None, :]
    idx, jdx = torch.triu_indices(4, 4)
    return A[..., idx, jdx]


def rotation_10d_to_quaternion(vec):
    A = _sym4(vec)
    return torch.linalg.eigh(A).eigenvectors[..., 0]


def quaternion_to_rotation_10d(q):
    A = torch.eye(4, dtype=q.dtype, device=q.device) - q[..., :, None] * q[..., None, :]
    idx, jdx = torch.triu_indices(4, 4)
    return A[..., idx, jdx]


def _se3_V(log_rot, eps=1e-4):
    nrms = (log_rot * log_rot).sum(-1)
    theta = nrms.clamp_min(eps).sqrt()
    K = _hat(log_rot)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)
    f1 = ((1 - theta.cos()) / theta**2)[..., None, None]
    f2 = ((theta - theta.sin()) / theta**3)[..., None, None]
    return eye + f1 * K + f2 * (K @ K)


# ----------------------------------------------------------------------------- SE(3)
def make_matrix(R, t):
    M = torch.zeros(R.shape[:-2] + (4, 4), dtype=R.dtype, device=R.device)
    M[..., :3, :3] = R
    M[..., :3, 3] = t
    M[..., 3, 3] = 1.0
    return M


def pose_from_params(rot, xyz, parameterization, convention=None, degrees=False):
    """diffdrr.pose.convert(rot, xyz, parameterization=, convention=, degrees=) -> (B,4,4)."""
    if parameterization == "euler_angles":
        if degrees:
            rot = torch.deg2rad(rot)
        R = euler_angles_to_matrix(rot, convention)
    elif parameterization == "axis_angle":
        R = axis_angle_to_matrix(rot)
    elif parameterization == "so3_log_map":
        R = so3_exp_map(rot)
    elif parameterization == "se3_log_map":
        R = so3_exp_map(rot)
        xyz = (_se3_V(rot) @ xyz[..., None])[..., 0]
    elif parameterization == "quaternion":
        R = quaternion_to_matrix(rot)
    elif parameterization == "rotation_6d":
        R = rotation_6d_to_matrix(rot)
    elif parameterization == "rotation_10d":
        R = quaternion_to_matrix(rotation_10d_to_quaternion(rot))
    elif parameterization == "quaternion_adjugate":
        R = quaternion_to_matrix(quaternion_adjugate_to_quaternion(rot))
    else:
        raise ValueError(parameterization)
    if knobs.CONVERT_TRANSLATION_IN_ROTATED_FRAME and parameterization != "se3_log_map":
        xyz = (R @ xyz[..., None])[..., 0]
    return make_matrix(R, xyz)


def params_from_pose(M, parameterization, convention=None, degrees=False):
    """RigidTransform.convert(parameterization, convention) -> (rot, xyz)."""
    R, t = M[..., :3, :3], M[..., :3, 3]
    if knobs.CONVERT_TRANSLATION_IN_ROTATED_FRAME and parameterization != "se3_log_map":
        t = (R.transpose(-1, -2) @ t[..., None])[..., 0]
    if parameterization == "euler_angles":
        rot = matrix_to_euler_angles(R, convention)
        if degrees:
            rot = torch.rad2deg(rot)
    elif parameterization == "axis_angle":
        rot = matrix_to_axis_angle(R)
    elif parameterization == "so3_log_map":
        rot = so3_log_map(R)
    elif parameterization == "se3_log_map":
        rot = so3_log_map(R)
        t = torch.linalg.solve(_se3_V(rot), t[..., None])[..., 0]
    elif parameterization == "quaternion":
        rot = matrix_to_quaternion(R)
    elif parameterization == "rotation_6d":
        rot = matrix_to_rotation_6d(R)
    elif parameterization == "rotation_10d":
        rot = quaternion_to_rotation_10d(matrix_to_quaternion(R))
    elif parameterization == "quaternion_adjugate":
        rot = quaternion_to_quaternion_adjugate(matrix_to_quaternion(R))
    else:
        raise ValueError(parameterization)
    return rot, t


def compose(A, B):
    """A.compose(B): apply A first, then B."""
    return B @ A if knobs.COMPOSE_APPLIES_SELF_FIRST else A @ B


def invert(M):
    R, t = M[..., :3, :3], M[..., :3, 3]
    Rt = R.transpose(-1, -2)
    return make_matrix(Rt, -(Rt @ t[..., None])[..., 0])


def apply(M, pts):
    """x' = M[:3,:3] x + M[:3,3] on (B,N,3) points (einsum 'bij,bnj->bni' in homogeneous form)."""
    ones = torch.ones_like(pts[..., :1])
    hom = torch.cat([pts, ones], -1)
    return torch.einsum("bij,bnj->bni", M.expand(pts.shape[0] if M.shape[0] == 1 else M.shape[0], 4, 4), hom)[..., :3]


_ = math  # keep import for users of this module
