"""Uniform 6-DoF pose sampling in the semantics of /root/reference/src/xvr/model/sampler.py:5-38:
ZXY Euler angles in degrees (circle-shifted to [-180, 180)), translations in mm, CPU ``torch.rand``."""

import torch

from .pose import convert

__all__ = ["get_random_pose", "random_pose_params"]


def _uniform(low, high, n, generator=None, circle_shift=False):
    x = (high - low) * torch.rand(n, 1, generator=generator) + low
    return ((x + 180) % 360) - 180 if circle_shift else x


def random_pose_params(alphamin, alphamax, betamin, betamax, gammamin, gammamax, txmin, txmax, tymin, tymax, tzmin,
                       tzmax, batch_size, generator=None):
    """(rot (B,3) degrees, xyz (B,3) mm) in the draw order of the reference sampler."""
    a = _uniform(alphamin, alphamax, batch_size, generator, True)
    b = _uniform(betamin, betamax, batch_size, generator, True)
    g = _uniform(gammamin, gammamax, batch_size, generator, True)
    tx = _uniform(txmin, txmax, batch_size, generator)
    ty = _uniform(tymin, tymax, batch_size, generator)
    tz = _uniform(tzmin, tzmax, batch_size, generator)
    return torch.cat([a, b, g], 1), torch.cat([tx, ty, tz], 1)


def get_random_pose(alphamin, alphamax, betamin, betamax, gammamin, gammamax, txmin, txmax, tymin, tymax, tzmin,
                    tzmax, batch_size, generator=None):
    rot, xyz = random_pose_params(alphamin, alphamax, betamin, betamax, gammamin, gammamax, txmin, txmax, tymin,
                                  tymax, tzmin, tzmax, batch_size, generator)
    return convert(rot, xyz, parameterization="euler_angles", convention="ZXY", degrees=True)
