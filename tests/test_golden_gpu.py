"""The CUDA path against the committed golden fixtures (tests/golden/oracle_v1.pt, made by make_golden.py)."""

import os

import pytest
import torch

import xvr_b200
from xvr_b200 import metrics
from xvr_b200.data import read
from xvr_b200.preprocess import XrayTransforms

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v1.pt"), weights_only=False)


def _drr(renderer, device):
    sub = read(GOLD["hu"], GOLD["labels"], affine=GOLD["affine"].numpy(), center_volume=False)
    d = GOLD["detector"]
    return xvr_b200.DRR(sub, d["sdd"], d["height"], d["delx"], d["width"], d["dely"], d["x0"], d["y0"],
                        reverse_x_axis=d["reverse_x_axis"], renderer=renderer).to(device)


@pytest.mark.parametrize("renderer", ["trilinear", "siddon"])
def test_render_and_gradients_match_golden(cuda, renderer):
    drr = _drr(renderer, cuda)
    g = GOLD[renderer]
    r, x = GOLD["rot"].to(cuda).requires_grad_(), GOLD["xyz"].to(cuda).requires_grad_()
    img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
    ref = g["img"].to(cuda)
    assert ((img - ref).norm() / ref.norm()).item() < 1e-4
    w = torch.linspace(0.5, 1.5, img.numel(), device=cuda).view_as(img)
    (img * w).sum().backward()
    # Siddon's pose gradient is a sum of voxel-value jumps at plane crossings: one-ulp differences between the CPU
    # that made the golden file and this GPU (sin/cos, matmul order) re-route a few of the 480 rays per pose
    # through other voxels, and the ORACLE ITSELF moves by ~2 % between the two devices on this coarse 32^3 case
    # (on the same device kernel and oracle agree to 1e-6 per ray, tests/test_siddon_gpu.py).
    tol = 2e-3 if renderer == "trilinear" else 5e-2
    assert ((r.grad.cpu() - g["grad_rot"]).norm() / g["grad_rot"].norm()).item() < tol
    assert ((x.grad.cpu() - g["grad_xyz"]).norm() / g["grad_xyz"].norm()).item() < tol
    pose = xvr_b200.convert(GOLD["rot"].to(cuda), GOLD["xyz"].to(cuda), parameterization="euler_angles", convention="ZXY")
    ch = drr(pose, mask_to_channels=True)
    refc = g["img_channels"].to(cuda)
    assert ch.shape == refc.shape
    assert ((ch - refc).norm() / refc.norm()).item() < 1e-3
    assert ((ch.sum(1) - refc.sum(1)).norm() / refc.sum(1).norm()).item() < 1e-4


def test_metrics_match_golden(cuda):
    m = GOLD["metrics"]
    x1, x2 = m["x1"].to(cuda), m["x2"].to(cuda)
    assert torch.allclose(metrics.NormalizedCrossCorrelation2d()(x1, x2).cpu(), m["ncc"], atol=1e-4)
    assert torch.allclose(metrics.NormalizedCrossCorrelation2d(9)(x1, x2).cpu(), m["ncc9"], atol=1e-4)
    assert torch.allclose(metrics.MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])(x1, x2).cpu(),
                          m["mncc"], atol=1e-4)
    assert torch.allclose(metrics.GradientNormalizedCrossCorrelation2d(11, 0.0).cuda()(x1, x2).cpu(), m["gncc11"],
                          atol=1e-4)
    # the transform chain that produced x1 from the golden DRR
    t = XrayTransforms(24, 20)(GOLD["trilinear"]["img"].to(cuda))
    assert torch.allclose(t.cpu(), m["x1"], atol=1e-5)
