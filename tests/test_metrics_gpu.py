"""Parity of the CUDA similarity kernels (NCC / multiscale NCC / gradient NCC / Sobel) against the oracle."""

import pytest
import torch

import oracle
from xvr_b200 import metrics

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _images(b, c, h, w, seed, device):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    base = torch.exp(-(xx**2 + yy**2) / 0.3)[None, None]
    x1 = (base * (1 + 0.3 * torch.rand(b, c, 1, 1, generator=g)) + 0.05 * torch.randn(b, c, h, w, generator=g) - 0.15) / 0.1
    x2 = (base.roll(3, -1) * 0.8 + 0.05 * torch.randn(b, c, h, w, generator=g) - 0.15) / 0.1
    # a perfectly flat region exercises the eps path of the patch variance
    x1[..., : h // 4, : w // 4] = -1.5
    x2[..., : h // 4, : w // 5] = -1.5
    return x1.to(device), x2.to(device)


def _check(fn_ours, fn_ref, x1, x2, grad_tol=5e-4):
    a1, a2 = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    b1, b2 = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    s, r = fn_ours(a1, a2), fn_ref(b1, b2)
    assert s.shape == r.shape == (x1.shape[0],)
    assert (s - r).abs().max().item() < TOL
    w = torch.linspace(0.5, 1.5, x1.shape[0], device=x1.device)
    (s * w).sum().backward()
    (r * w).sum().backward()
    for g, h in ((a1.grad, b1.grad), (a2.grad, b2.grad)):
        assert ((g - h).norm() / h.norm()).item() < grad_tol


@pytest.mark.parametrize("shape", [(3, 1, 64, 64), (2, 2, 37, 53)])
@pytest.mark.parametrize("patch", [None, 9, 11, 4])
def test_ncc(cuda, shape, patch):
    x1, x2 = _images(*shape, seed=0, device=cuda)
    _check(metrics.NormalizedCrossCorrelation2d(patch), lambda a, b: oracle.ncc(a, b, patch), x1, x2)


def test_multiscale_ncc_xvr_config(cuda):
    x1, x2 = _images(4, 1, 128, 128, seed=1, device=cuda)
    sim = metrics.MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])
    _check(sim, lambda a, b: oracle.multiscale_ncc(a, b, (None, 9), (0.5, 0.5)), x1, x2)


def test_gradient_ncc_xvr_config(cuda):
    x1, x2 = _images(3, 1, 96, 80, seed=2, device=cuda)
    sim = metrics.GradientNormalizedCrossCorrelation2d(patch_size=11, sigma=0.0).cuda()
    _check(sim, lambda a, b: oracle.gradient_ncc(a, b, 11, 0.0), x1, x2)


def test_gradient_ncc_with_blur(cuda):
    x1, x2 = _images(2, 1, 64, 64, seed=3, device=cuda)
    sim = metrics.GradientNormalizedCrossCorrelation2d(patch_size=None, sigma=1.5).cuda()
    _check(sim, lambda a, b: oracle.gradient_ncc(a, b, None, 1.5), x1, x2)


def test_sobel(cuda):
    x1, _ = _images(2, 1, 33, 47, seed=4, device=cuda)
    a, b = x1.clone().requires_grad_(), x1.clone().requires_grad_()
    s, r = metrics.Sobel(0.0).cuda()(a), oracle.sobel(b)
    assert torch.allclose(s, r, atol=1e-5)
    w = torch.rand_like(s)
    (s * w).sum().backward()
    (r * w).sum().backward()
    assert torch.allclose(a.grad, b.grad, atol=1e-4)


def test_identical_images_score_one(cuda):
    x1, _ = _images(2, 1, 64, 64, seed=5, device=cuda)
    x1 = x1 + 0.3 * torch.randn_like(x1)  # no flat windows
    s = metrics.MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])(x1, x1.clone())
    assert torch.allclose(s, torch.ones_like(s), atol=1e-3)


def test_double_geodesic(cuda):
    import xvr_b200
    from tests._scene import pose_params

    r1, x1 = pose_params(5, seed=1)
    r2, x2 = pose_params(5, seed=2)
    p1 = xvr_b200.convert(r1, x1, parameterization="euler_angles", convention="ZXY")
    p2 = xvr_b200.convert(r2, x2, parameterization="euler_angles", convention="ZXY")
    ours = metrics.DoubleGeodesicSE3(1020.0)(p1, p2)
    ref = oracle.double_geodesic(p1.matrix, p2.matrix, 1020.0)
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-4)


def test_hu_to_density_kernel_matches_oracle(cuda):
    from xvr_b200.data import synthetic_ct, transform_hu_to_density

    hu, _, _ = synthetic_ct(48, seed=5, device=cuda)
    for m in (1.0, 3.7, 10.0, 0.25):
        ours = transform_hu_to_density(hu, m)
        ref = oracle.hu_to_density(hu, m)
        assert torch.equal(ours, ref), m
    hu2 = hu.clone()
    hu2[hu2 > 350] = 100.0  # no bone at all
    assert torch.equal(transform_hu_to_density(hu2, 5.0), oracle.hu_to_density(hu2, 5.0))
    hu2.add_(50.0)  # in-place change must invalidate the cached statistics
    assert torch.equal(transform_hu_to_density(hu2, 2.0), oracle.hu_to_density(hu2, 2.0))
