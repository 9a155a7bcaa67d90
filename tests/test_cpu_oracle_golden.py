"""CPU suite, part 1: the oracle against its committed golden vectors, and its own self-consistency
(finite differences in fp64, DRR.__call__ == the trainer.py:283-289 decomposition)."""

import os

import numpy as np
import pytest
import torch

import oracle
from oracle import knobs

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v1.pt"), weights_only=False)


def _affinv():
    return torch.as_tensor(np.linalg.inv(GOLD["affine"].numpy()), dtype=torch.float32)[None]


def test_knobs_unchanged_since_golden():
    for k, v in GOLD["knobs"].items():
        assert getattr(knobs, k) == v, f"oracle knob {k} changed: regenerate tests/golden with make_golden.py"


def test_hu_to_density_golden():
    assert torch.equal(oracle.hu_to_density(GOLD["hu"], 1.0), GOLD["density"])


@pytest.mark.parametrize("renderer", ["trilinear", "siddon"])
def test_render_golden(renderer):
    torch.set_num_threads(1)
    r, x = GOLD["rot"].clone().requires_grad_(), GOLD["xyz"].clone().requires_grad_()
    pose = oracle.pose_from_params(r, x, "euler_angles", "ZXY")
    img = oracle.drr_forward(GOLD["density"], _affinv(), pose, reorient=oracle.REORIENT["AP"], renderer=renderer,
                             **GOLD["detector"])
    g = GOLD[renderer]
    assert torch.allclose(img, g["img"], rtol=1e-5, atol=1e-5)
    w = torch.linspace(0.5, 1.5, img.numel()).view_as(img)
    (img * w).sum().backward()
    assert torch.allclose(r.grad, g["grad_rot"], rtol=1e-3, atol=1e-2)
    assert torch.allclose(x.grad, g["grad_xyz"], rtol=1e-3, atol=1e-3)
    pose = oracle.pose_from_params(GOLD["rot"], GOLD["xyz"], "euler_angles", "ZXY")
    ch = oracle.drr_forward(GOLD["density"], _affinv(), pose, reorient=oracle.REORIENT["AP"], renderer=renderer,
                            mask=GOLD["labels"], **GOLD["detector"])
    assert torch.allclose(ch, g["img_channels"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(ch.sum(1, keepdim=True), img.detach(), rtol=1e-4, atol=1e-4)


def test_metrics_golden():
    m = GOLD["metrics"]
    assert torch.allclose(oracle.ncc(m["x1"], m["x2"]), m["ncc"], atol=1e-6)
    assert torch.allclose(oracle.ncc(m["x1"], m["x2"], 9), m["ncc9"], atol=1e-6)
    assert torch.allclose(oracle.multiscale_ncc(m["x1"], m["x2"], (None, 9), (0.5, 0.5)), m["mncc"], atol=1e-6)
    assert torch.allclose(oracle.gradient_ncc(m["x1"], m["x2"], 11, 0.0), m["gncc11"], atol=1e-6)


def test_pose_golden():
    for name, p in GOLD["poses"].items():
        M = oracle.pose_from_params(p["rot"], p["xyz"], name, "ZXY" if name == "euler_angles" else None)
        assert torch.allclose(M, p["matrix"], atol=1e-5), name


def test_call_equals_trainer_decomposition():
    """drr(pose) == detector -> ray length -> affine_inverse -> renderer -> reshape (trainer.py:283-289)."""
    d = GOLD["detector"]
    pose = oracle.pose_from_params(GOLD["rot"], GOLD["xyz"], "euler_angles", "ZXY")
    src, tgt = oracle.detector_rays(pose, oracle.REORIENT["AP"], d["height"], d["width"], d["delx"], d["dely"],
                                    d["x0"], d["y0"], d["sdd"], d["reverse_x_axis"])
    raylen = (tgt - src).norm(dim=-1).unsqueeze(1)
    src, tgt = oracle.apply(_affinv(), src), oracle.apply(_affinv(), tgt)
    img = oracle.trilinear_render(GOLD["density"], src, tgt, raylen).view(3, 1, d["height"], d["width"])
    assert torch.allclose(img, GOLD["trilinear"]["img"], rtol=1e-5, atol=1e-5)


def test_trilinear_gradient_against_finite_differences_fp64():
    torch.manual_seed(0)
    ax = torch.linspace(-1, 1, 12, dtype=torch.float64)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = torch.exp(-(X**2 + Y**2 + Z**2) / 0.3)
    src = torch.tensor([[[-20.0, 5.0, 6.0]]], dtype=torch.float64)
    tgt = torch.tensor([[[30.0, 6.5, 4.0], [28.0, 2.0, 9.0]]], dtype=torch.float64)
    raylen = torch.ones(1, 1, 2, dtype=torch.float64)

    def f(s, t):
        return oracle.trilinear_render(vol, s, t, raylen, n_points=40).sum()

    s, t = src.clone().requires_grad_(), tgt.clone().requires_grad_()
    f(s, t).backward()
    h = 1e-6
    for idx in [(0, 0, 0), (0, 0, 1), (0, 0, 2)]:
        e = torch.zeros_like(src)
        e[idx] = h
        fd = (f(src + e, tgt) - f(src - e, tgt)) / (2 * h)
        assert abs(fd - s.grad[idx]) < 1e-5 * max(1.0, abs(fd))
    for idx in [(0, 0, 0), (0, 1, 1), (0, 1, 2)]:
        e = torch.zeros_like(tgt)
        e[idx] = h
        fd = (f(src, tgt + e) - f(src, tgt - e)) / (2 * h)
        assert abs(fd - t.grad[idx]) < 1e-5 * max(1.0, abs(fd))


def test_siddon_equals_exact_line_integral_of_constant_volume():
    """A constant volume integrates to (chord length / segment length) * value * raylen."""
    vol = torch.full((8, 10, 12), 2.0)
    src = torch.tensor([[[-5.0, 3.3, 4.1]]])
    tgt = torch.tensor([[[20.0, 6.2, 9.7]]])
    raylen = torch.full((1, 1, 1), 7.0)
    img = oracle.siddon_render(vol, src, tgt, raylen, voxel_shift=0.5)
    d = tgt - src
    lo = (torch.tensor([-0.5, -0.5, -0.5]) - src) / d
    hi = (torch.tensor([7.5, 9.5, 11.5]) - src) / d
    a0 = torch.minimum(lo, hi).max()
    a1 = torch.maximum(lo, hi).min()
    assert torch.allclose(img.flatten(), (a1 - a0) * 2.0 * 7.0, rtol=1e-5)


def test_siddon_segments_consistent_with_render():
    g = torch.Generator().manual_seed(0)
    vol = torch.rand(9, 8, 7, generator=g)
    src = torch.tensor([[[-6.0, 2.2, 3.3]], [[4.0, -9.0, 1.0]]])
    tgt = torch.stack([torch.tensor([[15.0, 5.0, 4.0], [14.0, 7.5, 1.0]]), torch.tensor([[4.5, 12.0, 5.0], [1.0, 15.0, 6.0]])])
    raylen = torch.ones(2, 1, 2)
    idx, seg = oracle.siddon_segments(vol.shape, src, tgt)
    vals = torch.where(idx >= 0, vol.flatten()[idx.clamp_min(0)], torch.zeros(()))
    assert torch.allclose((vals * seg).sum(-1), oracle.siddon_render(vol, src, tgt, raylen)[:, 0], atol=1e-6)
