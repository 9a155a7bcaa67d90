"""SE(3) pose parameterisations -- the host-side mirror of ``diffdrr.pose``.

Interface pinned by the xvr call sites: ``convert(rot, xyz, parameterization=, convention=, degrees=)``
(/root/reference/src/xvr/model/sampler.py:29-31, model/network.py:49-54, model/trainer.py:335-337),
``RigidTransform`` with ``.matrix``, ``.compose``, ``.inverse``, ``.convert``, ``__getitem__``, ``__len__``,
``__matmul__`` and ``__call__(points)`` (model/trainer.py:193,204,270,289; model/loss.py:45-49;
registrar/base.py:168,201,264; metrics/evaluator.py:29) and ``make_matrix``.

These are O(B) operations on 4x4 matrices -- not the bandwidth path.  On CUDA tensors ``convert`` is ONE launch
each way (csrc/pose.cu: forward per pose, backward per (pose, parameter) by forward-mode differentiation) for every
parameterisation with a closed form; the tensor-op formulation below is the specification those kernels are tested
against (and what CPU tensors and ``rotation_10d`` run).
"""

import ctypes
import os

import torch

from . import _conventions as conv

__all__ = ["RigidTransform", "convert", "make_matrix", "N_ANGULAR_COMPONENTS"]

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


# --------------------------------------------------------------------------------------------- rotations
def _elementary(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    o, z = torch.ones_like(angle), torch.zeros_like(angle)
    rows = {
        "X": (o, z, z, z, c, -s, z, s, c),
        "Y": (c, z, s, z, o, z, -s, z, c),
        "Z": (c, -s, z, s, c, z, z, z, o),
    }[axis]
    return torch.stack(rows, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(angles, convention):
    if len(convention) != 3 or any(c not in "XYZ" for c in convention):
        raise ValueError(f"Invalid Euler convention {convention!r}")
    r0, r1, r2 = (_elementary(c, a) for c, a in zip(convention, angles.unbind(-1)))
    return r0 @ r1 @ r2


def _angle_from_tan(axis, other, data, horizontal, tait_bryan):
    i1, i2 = {"X": (2, 1), "Y": (0, 2), "Z": (1, 0)}[axis]
    if horizontal:
        i2, i1 = i1, i2
    even = (axis + other) in ("XY", "YZ", "ZX")
    if horizontal == even:
        return torch.atan2(data[..., i1], data[..., i2])
    if tait_bryan:
        return torch.atan2(-data[..., i2], data[..., i1])
    return torch.atan2(data[..., i2], -data[..., i1])


def matrix_to_euler_angles(R, convention):
    ix = {"X": 0, "Y": 1, "Z": 2}
    i0, i2 = ix[convention[0]], ix[convention[2]]
    tait_bryan = i0 != i2
    if tait_bryan:
        central = torch.asin(R[..., i0, i2] * (-1.0 if i0 - i2 in (-1, 2) else 1.0))
    else:
        central = torch.acos(R[..., i0, i0])
    return torch.stack(
        (
            _angle_from_tan(convention[0], convention[1], R[..., i2], False, tait_bryan),
            central,
            _angle_from_tan(convention[2], convention[1], R[..., i0, :], True, tait_bryan),
        ),
        -1,
    )


def quaternion_to_matrix(q):
    w, x, y, z = q.unbind(-1)
    k = 2.0 / (q * q).sum(-1)
    m = (
        1 - k * (y * y + z * z), k * (x * y - z * w), k * (x * z + y * w),
        k * (x * y + z * w), 1 - k * (x * x + z * z), k * (y * z - x * w),
        k * (x * z - y * w), k * (y * z + x * w), 1 - k * (x * x + y * y),
    )
    return torch.stack(m, -1).reshape(q.shape[:-1] + (3, 3))


def matrix_to_quaternion(R):
    lead = R.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = R.reshape(lead + (9,)).unbind(-1)
    sq = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    q_abs = torch.where(sq > 0, sq.clamp_min(1e-30).sqrt(), torch.zeros_like(sq))
    rows = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    ) / (2.0 * q_abs[..., None].clamp_min(0.1))
    pick = q_abs.argmax(-1)
    q = torch.gather(rows, -2, pick[..., None, None].expand(lead + (1, 4)))[..., 0, :]
    return torch.where(q[..., :1] < 0, -q, q)


def _hat(v):
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack((o, -z, y, z, o, -x, -y, x, o), -1).reshape(v.shape[:-1] + (3, 3))


def so3_exp_map(w, eps=1e-4):
    theta = (w * w).sum(-1).clamp_min(eps).sqrt()
    K = _hat(w)
    eye = torch.eye(3, dtype=w.dtype, device=w.device)
    a = (theta.sin() / theta)[..., None, None]
    b = ((1 - theta.cos()) / (theta * theta))[..., None, None]
    return eye + a * K + b * (K @ K)


def so3_log_map(R, eps=1e-4):
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    phi = torch.acos(((tr - 1.0) * 0.5).clamp(-1.0 + 1e-7, 1.0 - 1e-7))
    sin = phi.sin()
    fac = torch.where(sin.abs() > 0.5 * eps, phi / (2.0 * sin.clamp_min(1e-12)), 0.5 + phi * phi / 12.0)
    skew = fac[..., None, None] * (R - R.transpose(-1, -2))
    return torch.stack((skew[..., 2, 1], skew[..., 0, 2], skew[..., 1, 0]), -1)


def _se3_V(w, eps=1e-4):
    theta = (w * w).sum(-1).clamp_min(eps).sqrt()
    K = _hat(w)
    eye = torch.eye(3, dtype=w.dtype, device=w.device)
    a = ((1 - theta.cos()) / theta**2)[..., None, None]
    b = ((theta - theta.sin()) / theta**3)[..., None, None]
    return eye + a * K + b * (K @ K)


def axis_angle_to_matrix(aa):
    return so3_exp_map(aa, eps=1e-12)


def matrix_to_axis_angle(R):
    q = matrix_to_quaternion(R)
    n = q[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return q[..., 1:] / s


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), -2)


# Index tables of the symmetric 4x4 <-> 10-vector maps, cached per device: building them on the fly would copy a
# CPU index tensor to the GPU on every call, which a CUDA-graph capture (the graphed training step) forbids.
_SYM4_GATHER = (0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9)  # flat 4x4 position -> component of the 10-vector
_TRIU4_GATHER = (0, 1, 2, 3, 5, 6, 7, 10, 11, 15)                # component -> flat 4x4 position (row-major triu)
_index_cache = {}


def _const_index(name, values, device):
    key = (name, device)
    if key not in _index_cache:
        _index_cache[key] = torch.tensor(values, dtype=torch.long, device=device)
    return _index_cache[key]


def _sym4(v):
    """10-vector (row-major upper triangle) -> symmetric 4x4."""
    return v.index_select(-1, _const_index("sym4", _SYM4_GATHER, v.device)).reshape(v.shape[:-1] + (4, 4))


def _triu4(A):
    """4x4 -> its row-major upper triangle as a 10-vector."""
    return A.reshape(A.shape[:-2] + (16,)).index_select(-1, _const_index("triu4", _TRIU4_GATHER, A.device))


def quaternion_adjugate_to_quaternion(v):
    A = _sym4(v)
    n = A.norm(dim=-2)
    pick = n.argmax(-1)
    col = torch.gather(A, -1, pick[..., None, None].expand(A.shape[:-1] + (1,)))[..., 0]
    return col / n.gather(-1, pick[..., None])


def rotation_10d_to_quaternion(v):
    return torch.linalg.eigh(_sym4(v)).eigenvectors[..., 0]


# --------------------------------------------------------------------------------------------- SE(3)
def make_matrix(R, t):
    """Assemble (B,4,4) from a rotation (B,3,3) and translation (B,3)."""
    top = torch.cat([R, t[..., None]], -1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom[..., 0, 3] = 1.0
    return torch.cat([top, bottom], -2)


def _rotation(rot, xyz, parameterization, convention, degrees):
    if parameterization == "euler_angles":
        if convention is None:
            raise ValueError("euler_angles needs a convention such as 'ZXY'")
        return euler_angles_to_matrix(torch.deg2rad(rot) if degrees else rot, convention), xyz
    if parameterization == "axis_angle":
        return axis_angle_to_matrix(rot), xyz
    if parameterization == "so3_log_map":
        return so3_exp_map(rot), xyz
    if parameterization == "se3_log_map":
        return so3_exp_map(rot), (_se3_V(rot) @ xyz[..., None])[..., 0]
    if parameterization == "quaternion":
        return quaternion_to_matrix(rot), xyz
    if parameterization == "rotation_6d":
        return rotation_6d_to_matrix(rot), xyz
    if parameterization == "rotation_10d":
        return quaternion_to_matrix(rotation_10d_to_quaternion(rot)), xyz
    if parameterization == "quaternion_adjugate":
        return quaternion_to_matrix(quaternion_adjugate_to_quaternion(rot)), xyz
    raise ValueError(f"Unknown parameterization {parameterization!r}; choose from {list(N_ANGULAR_COMPONENTS)}")


# parameterisations of csrc/pose.cu (enum PoseKind)
POSE_KERNEL_KINDS = {"euler_angles": 0, "axis_angle": 1, "so3_log_map": 2, "se3_log_map": 3, "quaternion": 4,
                     "rotation_6d": 5, "quaternion_adjugate": 6}


def _euler_axes(convention):
    if convention is None or len(convention) != 3 or any(c not in "XYZ" for c in convention):
        raise ValueError(f"Invalid Euler convention {convention!r}")
    return (ctypes.c_int * 3)(*("XYZ".index(c) for c in convention))


class _PoseKernel(torch.autograd.Function):
    """(rot (B,n), xyz (B,3)) -> pose (B,4,4) [and, with camera constants, cam2vox / cam2world (B,3,4)] through
    xvr_pose_fwd / xvr_pose_bwd: one launch each way."""

    @staticmethod
    def forward(ctx, rot, xyz, kind, convention, degrees, camera):
        from ._lib import call, cuda_f32, ptr, stream  # noqa: PLC0415

        rot, xyz = cuda_f32(rot, "rotation"), cuda_f32(xyz, "translation")
        B, n_rot = rot.shape
        axes = _euler_axes(convention) if kind == 0 else None
        scale = torch.pi / 180.0 if (degrees and kind == 0) else 1.0
        consts = (B, kind, n_rot, axes, int(conv.CONVERT_TRANSLATION_IN_ROTATED_FRAME), float(scale))
        ctx.save_for_backward(rot, xyz)
        if camera is None:
            pose = torch.empty(B, 4, 4, device=rot.device, dtype=torch.float32)
            ctx.consts = (*consts, None, None)
            if B > 0:
                call("xvr_pose_fwd", ptr(rot), ptr(xyz), *ctx.consts, ptr(pose), None, None, stream())
            return pose
        reorient16, affinv16 = camera
        ctx.consts = (*consts, (ctypes.c_float * 16)(*reorient16), (ctypes.c_float * 16)(*affinv16))
        cam2world = torch.empty(B, 3, 4, device=rot.device, dtype=torch.float32)
        cam2vox = torch.empty(B, 3, 4, device=rot.device, dtype=torch.float32)
        if B > 0:
            call("xvr_pose_fwd", ptr(rot), ptr(xyz), *ctx.consts, None, ptr(cam2world), ptr(cam2vox), stream())
        ctx.mark_non_differentiable(cam2world)  # only feeds the ray length, which a rigid motion leaves unchanged
        return cam2vox, cam2world

    @staticmethod
    def backward(ctx, g, *_unused):
        from ._lib import call, cuda_f32, ptr, stream  # noqa: PLC0415

        rot, xyz = ctx.saved_tensors
        grot, gxyz = torch.empty_like(rot), torch.empty_like(xyz)
        if rot.shape[0] > 0:
            g = cuda_f32(g, "grad")
            camera = ctx.consts[-1] is not None
            call("xvr_pose_bwd", ptr(rot), ptr(xyz), *ctx.consts, None if camera else ptr(g), ptr(g) if camera else None,
                 ptr(grot), ptr(gxyz), stream())
        return grot, gxyz, None, None, None, None


def _kernel_eligible(rot, xyz, parameterization):
    return (parameterization in POSE_KERNEL_KINDS and torch.is_tensor(rot) and rot.is_cuda and xyz.is_cuda
            and rot.dim() == 2 and xyz.dim() == 2 and rot.shape[0] == xyz.shape[0] and xyz.shape[1] == 3
            and rot.shape[1] == N_ANGULAR_COMPONENTS[parameterization] and rot.dtype == torch.float32
            and xyz.dtype == torch.float32 and os.environ.get("XVR_B200_FUSED_POSE", "1") == "1")


def convert(*args, parameterization, convention=None, degrees=False):
    """``convert(rot, xyz, parameterization=..., convention=..., degrees=...) -> RigidTransform``."""
    if len(args) != 2:
        raise TypeError("convert(rot, xyz, parameterization=..., convention=...)")
    rot, xyz = args
    if _kernel_eligible(rot, xyz, parameterization):
        return RigidTransform(_PoseKernel.apply(rot, xyz, POSE_KERNEL_KINDS[parameterization], convention, degrees,
                                                None))
    R, t = _rotation(rot, xyz, parameterization, convention, degrees)
    if conv.CONVERT_TRANSLATION_IN_ROTATED_FRAME and parameterization != "se3_log_map":
        t = (R @ t[..., None])[..., 0]
    return RigidTransform(make_matrix(R, t))


class RigidTransform(torch.nn.Module):
    """Batched rigid transform acting on column vectors: x' = M [x; 1]."""

    def __init__(self, matrix):
        super().__init__()
        if matrix.dim() == 2:
            matrix = matrix[None]
        if matrix.shape[-2:] != (4, 4):
            raise ValueError(f"expected (B,4,4), got {tuple(matrix.shape)}")
        self.register_buffer("_matrix", matrix)

    @property
    def matrix(self):
        return self._matrix

    @property
    def rotation(self):
        return self._matrix[..., :3, :3]

    @property
    def translation(self):
        return self._matrix[..., :3, 3]

    def __len__(self):
        return len(self._matrix)

    def __getitem__(self, idx):
        return RigidTransform(self._matrix[idx])

    def forward(self, x):
        """Apply to points (B,N,3) (batch of 1 broadcasts)."""
        M = self._matrix.to(x.dtype)
        return x @ M[..., :3, :3].transpose(-1, -2) + M[..., None, :3, 3]

    def inverse(self):
        Rt = self.rotation.transpose(-1, -2)
        return RigidTransform(make_matrix(Rt, -(Rt @ self.translation[..., None])[..., 0]))

    def compose(self, other):
        """``A.compose(B)``: apply A, then B."""
        if conv.COMPOSE_APPLIES_SELF_FIRST:
            return RigidTransform(other.matrix @ self.matrix)
        return RigidTransform(self.matrix @ other.matrix)

    def __matmul__(self, other):
        return RigidTransform(self.matrix @ other.matrix)

    def convert(self, parameterization, convention=None, degrees=False):
        """Inverse of :func:`convert`: ``-> (rot, xyz)``."""
        R, t = self.rotation, self.translation
        if conv.CONVERT_TRANSLATION_IN_ROTATED_FRAME and parameterization != "se3_log_map":
            t = (R.transpose(-1, -2) @ t[..., None])[..., 0]
        if parameterization == "euler_angles":
            rot = matrix_to_euler_angles(R, convention)
            rot = torch.rad2deg(rot) if degrees else rot
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
            rot = R[..., :2, :].clone().reshape(R.shape[:-2] + (6,))
        elif parameterization == "rotation_10d":
            q = matrix_to_quaternion(R)
            rot = _triu4(torch.eye(4, dtype=q.dtype, device=q.device) - q[..., :, None] * q[..., None, :])
        elif parameterization == "quaternion_adjugate":
            q = matrix_to_quaternion(R)
            rot = _triu4(q[..., :, None] * q[..., None, :])
        else:
            raise ValueError(f"Unknown parameterization {parameterization!r}")
        return rot, t
