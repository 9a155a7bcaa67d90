"""Oracle restatement of diffdrr.metrics (image similarity + SE(3) geodesics) -- TEST INFRA ONLY.

Follows diffdrr 0.6.0 ``diffdrr/metrics.py`` (NormalizedCrossCorrelation2d, to_patches,
MultiscaleNormalizedCrossCorrelation2d, GradientNormalizedCrossCorrelation2d, Sobel,
DoubleGeodesicSE3) as pinned by the xvr call sites /root/reference/src/xvr/model/loss.py:16-29 and
/root/reference/src/xvr/registrar/base.py:115-123.  PARITY UNPINNED (see oracle/__init__.py).
"""

import torch
import torch.nn.functional as F

from . import knobs
from .pose import so3_log_map

__all__ = ["ncc", "multiscale_ncc", "sobel", "gradient_ncc", "double_geodesic"]


def _zscore(x, eps):
    mu = x.mean(dim=[-1, -2], keepdim=True)
    var = x.var(dim=[-1, -2], keepdim=True, correction=0) + eps
    return (x - mu) / var.sqrt()


def _to_patches(x, p):
    # unfold -> (b, c, nH, nW, p, p); the window is normalised over its own p x p extent
    x = x.unfold(2, p, 1).unfold(3, p, 1).contiguous()
    b, c, nh, nw, p1, p2 = x.shape
    return x.reshape(b, c * nh * nw, p1, p2)


def ncc(x1, x2, patch_size=None, eps=None):
    """NormalizedCrossCorrelation2d(patch_size, eps)(x1, x2) -> (B,)."""
    eps = knobs.NCC_EPS if eps is None else eps
    if patch_size is not None:
        x1, x2 = _to_patches(x1, patch_size), _to_patches(x2, patch_size)
    _, c, h, w = x1.shape
    z1, z2 = _zscore(x1, eps), _zscore(x2, eps)
    return torch.einsum("b...,b...->b", z1, z2) / (c * h * w)


def multiscale_ncc(x1, x2, patch_sizes=(None, 9), patch_weights=(0.5, 0.5), eps=None):
    """MultiscaleNormalizedCrossCorrelation2d(patch_sizes, patch_weights)(x1, x2) -> (B,)."""
    total = 0.0
    for p, w in zip(patch_sizes, patch_weights):
        total = total + w * ncc(x1, x2, p, eps)
    return total


def sobel(img, sigma=0.0):
    """Sobel: bias-free Conv2d(1, 2, 3, padding=1) with fixed Gx/Gy; Gaussian blur only if sigma>0."""
    gx = torch.tensor([[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]])
    gy = torch.tensor([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]])
    G = torch.stack([gx, gy]).unsqueeze(1).to(img)
    if sigma > 0:
        k = torch.arange(5, dtype=img.dtype, device=img.device) - 2
        g1 = torch.exp(-0.5 * (k / sigma) ** 2)
        g1 = g1 / g1.sum()
        pad = F.pad(img, (2, 2, 2, 2), mode="reflect")
        img = F.conv2d(F.conv2d(pad, g1.view(1, 1, 1, 5)), g1.view(1, 1, 5, 1))
    return F.conv2d(img, G, padding=1)


def gradient_ncc(x1, x2, patch_size=None, sigma=0.0, eps=None):
    """GradientNormalizedCrossCorrelation2d(patch_size, sigma)(x1, x2) -> (B,)."""
    return ncc(sobel(x1, sigma), sobel(x2, sigma), patch_size, eps)


def double_geodesic(pose_a, pose_b, sdd, eps=None):
    """DoubleGeodesicSE3(sdd, eps)(A, B) -> (angular, translational, double), each (B,)."""
    eps = knobs.GEODESIC_EPS if eps is None else eps
    Ra, Rb = pose_a[..., :3, :3], pose_b[..., :3, :3]
    ang = (sdd / 2.0) * so3_log_map(Ra.transpose(-1, -2) @ Rb).norm(dim=-1)
    tra = (pose_a[..., :3, 3] - pose_b[..., :3, 3]).norm(dim=-1)
    return ang, tra, (ang.square() + tra.square() + eps).sqrt()
