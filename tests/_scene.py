"""Shared scene builders for the parity tests: synthetic CT, DRR module, pose batches."""

import torch

import xvr_b200
from xvr_b200.data import read, synthetic_ct
from xvr_b200.sampler import random_pose_params

# pelvis-script pose ranges narrowed so that every pose keeps the volume in view
POSE_RANGES = dict(alphamin=-45, alphamax=45, betamin=-45, betamax=45, gammamin=-15, gammamax=15, txmin=-50,
                   txmax=50, tymin=700, tymax=900, tzmin=-50, tzmax=50)
SDD = 1020.0


def pixel_size(height):
    return 1.08821875 * 256.0 / height


def make_subject(n, seed=0, with_labels=False, multiplier=1.0):
    hu, lab, affine = synthetic_ct(n, seed=seed, with_labels=with_labels)
    return read(hu, lab, affine=affine, bone_attenuation_multiplier=multiplier)


def make_drr(n, height, renderer="trilinear", seed=0, with_labels=False, device="cuda", width=None, **kw):
    sub = make_subject(n, seed, with_labels)
    drr = xvr_b200.DRR(sub, SDD, height, pixel_size(height), width=width, renderer=renderer, reverse_x_axis=False, **kw)
    return drr.to(device)


def pose_params(batch, seed=0, device="cuda", radians=True):
    g = torch.Generator().manual_seed(seed)
    rot, xyz = random_pose_params(**POSE_RANGES, batch_size=batch, generator=g)
    if radians:
        rot = torch.deg2rad(rot)
    return rot.to(device), xyz.to(device)


def oracle_render(drr, rot, xyz, renderer="trilinear", mask=None, **kw):
    """The oracle's DRR.forward on the same module state (device follows the inputs)."""
    import oracle

    pose = oracle.pose_from_params(rot, xyz, "euler_angles", "ZXY")
    d = drr.detector
    return oracle.drr_forward(
        drr.density.to(rot.device), drr._affine_inverse.to(rot.device)[None], pose,
        reorient=d._reorient.to(rot.device), height=d.height, width=d.width, delx=d.delx, dely=d.dely, x0=d.x0,
        y0=d.y0, sdd=d.sdd, reverse_x_axis=d.reverse_x_axis, renderer=renderer, mask=mask, **kw)


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
