"""Rotation conversions of the pose layer against an INDEPENDENT implementation: scipy.spatial.transform.

DiffDRR's pose.py ports PyTorch3D's conversion functions; neither package is available here,
but the conventions they document are checkable against scipy: PyTorch3D's ``euler_angles_to_matrix(a, "ZXY")`` is
the intrinsic composition R_Z(a0) R_X(a1) R_Y(a2) (scipy's upper-case axis strings), its quaternions are real-first
(scipy: scalar-last), ``axis_angle`` / ``so3`` logarithms are rotation vectors, and the SE(3) exponential couples the
translation through the left Jacobian V (scipy >= 1.16 ``RigidTransform.from_exp_coords``).  Both the oracle and the
product are held to scipy, so this part of the oracle is pinned by a third party rather than by itself.
"""

import itertools

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

import oracle
import xvr_b200

CONVENTIONS = ["".join(p) for p in itertools.permutations("XYZ")] + ["XYX", "XZX", "YXY", "YZY", "ZXZ", "ZYZ"]


def _angles(n, seed):
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(n, 3, generator=g) - 0.5) * 2.0
    return a * torch.tensor([3.0, 1.4, 3.0])  # middle angle inside (-pi/2, pi/2): no gimbal lock for any convention


def _rot(T):
    return T.matrix[:, :3, :3].double().numpy()


@pytest.mark.parametrize("convention", CONVENTIONS)
def test_euler_angles_are_intrinsic_rotations_in_the_order_of_the_convention_string(convention):
    ang = _angles(16, 1)
    ref = Rotation.from_euler(convention, ang.double().numpy()).as_matrix()
    zero = torch.zeros(16, 3)
    ours = _rot(xvr_b200.convert(ang, zero, parameterization="euler_angles", convention=convention))
    orc = oracle.pose_from_params(ang, zero, "euler_angles", convention)[:, :3, :3].double().numpy()
    assert np.abs(ours - ref).max() < 2e-6
    assert np.abs(orc - ref).max() < 2e-6
    # degrees=True is the same rotation
    deg = _rot(xvr_b200.convert(torch.rad2deg(ang), zero, parameterization="euler_angles", convention=convention,
                                degrees=True))
    assert np.abs(deg - ref).max() < 5e-6


@pytest.mark.parametrize("convention", CONVENTIONS)
def test_matrix_to_euler_angles_inverts_scipy_rotations(convention):
    ref = Rotation.random(16, random_state=2)
    M = torch.eye(4).repeat(16, 1, 1)
    M[:, :3, :3] = torch.as_tensor(ref.as_matrix(), dtype=torch.float32)
    rot, _ = xvr_b200.RigidTransform(M).convert("euler_angles", convention)
    back = Rotation.from_euler(convention, rot.double().numpy()).as_matrix()
    assert np.abs(back - ref.as_matrix()).max() < 1e-5
    rot_o, _ = oracle.params_from_pose(M, "euler_angles", convention)
    assert np.abs(Rotation.from_euler(convention, rot_o.double().numpy()).as_matrix() - ref.as_matrix()).max() < 1e-5


def test_quaternions_are_real_first():
    ref = Rotation.random(32, random_state=3)
    xyzw = torch.as_tensor(ref.as_quat(), dtype=torch.float32)
    wxyz = torch.cat([xyzw[:, 3:], xyzw[:, :3]], dim=1)
    zero = torch.zeros(32, 3)
    assert np.abs(_rot(xvr_b200.convert(wxyz, zero, parameterization="quaternion")) - ref.as_matrix()).max() < 2e-6
    # unnormalised quaternions describe the same rotation (the CNN head emits them unnormalised)
    assert np.abs(_rot(xvr_b200.convert(3.0 * wxyz, zero, parameterization="quaternion")) - ref.as_matrix()).max() < 2e-6
    M = torch.eye(4).repeat(32, 1, 1)
    M[:, :3, :3] = torch.as_tensor(ref.as_matrix(), dtype=torch.float32)
    q, _ = xvr_b200.RigidTransform(M).convert("quaternion")
    q = q.double().numpy()
    back = Rotation.from_quat(np.concatenate([q[:, 1:], q[:, :1]], axis=1)).as_matrix()
    assert np.abs(back - ref.as_matrix()).max() < 1e-5


@pytest.mark.parametrize("name", ["axis_angle", "so3_log_map"])
def test_rotation_vectors(name):
    ref = Rotation.random(32, random_state=4)
    vec = torch.as_tensor(ref.as_rotvec(), dtype=torch.float32)
    vec[0] = 0.0           # identity
    vec[1] *= 1e-4 / vec[1].norm()  # tiny angle: the series branch
    ref = Rotation.from_rotvec(vec.double().numpy())
    zero = torch.zeros(32, 3)
    assert np.abs(_rot(xvr_b200.convert(vec, zero, parameterization=name)) - ref.as_matrix()).max() < 2e-6
    assert np.abs(oracle.pose_from_params(vec, zero, name)[:, :3, :3].double().numpy() - ref.as_matrix()).max() < 2e-6
    M = torch.eye(4).repeat(32, 1, 1)
    M[:, :3, :3] = torch.as_tensor(ref.as_matrix(), dtype=torch.float32)
    back, _ = xvr_b200.RigidTransform(M).convert(name)
    err = np.abs(Rotation.from_rotvec(back.double().numpy()).as_matrix() - ref.as_matrix()).reshape(32, -1).max(1)
    # PyTorch3D's so3_log_map takes acos of the trace and divides by sin: ill-conditioned within ~0.1 rad of pi in fp32
    # (3.5e-4 at 3.131 rad).  That is the reference formulation's own accuracy, not a convention.
    near_pi = vec.norm(dim=1).numpy() > 3.0
    assert err[~near_pi].max() < 1e-5
    assert near_pi.sum() == 0 or err[near_pi].max() < (2e-3 if name == "so3_log_map" else 1e-5)


def test_se3_exponential_couples_the_translation_through_the_left_jacobian():
    from scipy.spatial.transform import RigidTransform as ScipyRigid

    g = torch.Generator().manual_seed(5)
    w = (torch.rand(16, 3, generator=g) - 0.5) * 3.0
    v = (torch.rand(16, 3, generator=g) - 0.5) * 200.0
    ref = ScipyRigid.from_exp_coords(np.concatenate([w.double().numpy(), v.double().numpy()], axis=1)).as_matrix()
    ours = xvr_b200.convert(w, v, parameterization="se3_log_map").matrix.double().numpy()
    orc = oracle.pose_from_params(w, v, "se3_log_map").double().numpy()
    assert np.abs(ours[:, :3, :3] - ref[:, :3, :3]).max() < 2e-6
    assert np.abs(ours[:, :3, 3] - ref[:, :3, 3]).max() < 1e-3  # translations are O(100 mm) in fp32
    assert np.abs(orc - ours).max() < 1e-3
    w2, v2 = xvr_b200.RigidTransform(torch.as_tensor(ref, dtype=torch.float32)).convert("se3_log_map")
    assert torch.allclose(w2, w, atol=1e-4) and torch.allclose(v2, v, atol=2e-2)


def test_rotation_6d_is_gram_schmidt_on_two_rows():
    """Zhou et al. 2019 as PyTorch3D states it: orthonormalise the first two ROWS, third row = their cross product."""
    g = torch.Generator().manual_seed(6)
    d6 = torch.randn(16, 6, generator=g)
    a1, a2 = d6[:, :3].double().numpy(), d6[:, 3:].double().numpy()
    b1 = a1 / np.linalg.norm(a1, axis=1, keepdims=True)
    b2 = a2 - (b1 * a2).sum(1, keepdims=True) * b1
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    ref = np.stack([b1, b2, np.cross(b1, b2)], axis=1)
    ours = _rot(xvr_b200.convert(d6, torch.zeros(16, 3), parameterization="rotation_6d"))
    assert np.abs(ours - ref).max() < 2e-6
    assert np.abs(np.linalg.det(ours) - 1.0).max() < 1e-5
