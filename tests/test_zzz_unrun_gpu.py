"""GPU tests of code written AFTER round 1's GPU budget was spent -- never run on a device yet.

Each is a non-strict ``xfail`` (a pass is reported as XPASS, a failure does not fail the suite) and the file sorts
after every validated test, so that nothing here can mask or disturb them.  Round 2: run, fix what fails, move the
tests next to their validated neighbours and drop the markers.

  * ``xvr_regsim`` -- fused registration similarity, opt-in via ``Registrar(fused_similarity=True)``;
  * the gather (version 1) dL/dvolume kernel at the edge-pose set: measured 6 % off before its pixel window was
    replaced by the projected-corner box near the source; the brick kernel (default) passes.
"""

import pytest
import torch

import xvr_b200
from tests._scene import rel_l2
from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ, _volume_gradient_vs_oracle
from xvr_b200._lib import call

pytestmark = pytest.mark.gpu

_UNRUN = pytest.mark.xfail(strict=False, reason="written after round 1's GPU budget was spent; not yet run on a B200")


@_UNRUN
def test_gather_volume_gradient_edge_poses(cuda):
    from tests._scene import make_drr

    call("xvr_set_volgrad_version", 1)
    try:
        drr = make_drr(64, 32)
        _volume_gradient_vs_oracle(drr, torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda))
    finally:
        call("xvr_set_volgrad_version", 2)


@_UNRUN
@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (3, 40, 33)])
def test_fused_registration_similarity_matches_the_composition(cuda, B, H, W):
    """xvr_regsim (value + gradient, nine launches) against XrayTransforms -> beta mNCC + (1 - beta) GradNCC ->
    .sum() -> autograd built from the unfused modules, on images with ties at the minimum (DRR background) and at
    the maximum."""
    from xvr_b200.metrics import (GradientNormalizedCrossCorrelation2d, MultiscaleNormalizedCrossCorrelation2d,
                                  RegistrationSimilarity)
    from xvr_b200.preprocess import XrayTransforms

    g = torch.Generator().manual_seed(7)
    transform = XrayTransforms(H, W)
    fixed = transform((torch.rand(B, 1, H, W, generator=g) * 5.0).to(cuda))
    moving = (torch.rand(B, 1, H, W, generator=g) * 3.0).to(cuda)
    moving[:, :, :6, :7] = 0.0
    moving[0, 0, 10, 10] = moving[-1, 0, 20, 5] = 4.0
    beta = 0.3
    sim1 = MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])
    sim2 = GradientNormalizedCrossCorrelation2d(11, sigma=0.0).to(cuda)

    m1 = moving.clone().requires_grad_()
    y = transform(m1)
    ref = (beta * sim1(fixed, y) + (1 - beta) * sim2(fixed, y)).sum()
    ref.backward()

    m2 = moving.clone().requires_grad_()
    out = RegistrationSimilarity(fixed, 9, 11, beta=beta)(m2)
    (2.0 * out).backward()
    assert out.shape == ()
    assert abs(out.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert rel_l2(m2.grad, 2.0 * m1.grad) < 1e-4
    assert (m2.grad - 2.0 * m1.grad).abs().max().item() < 1e-4 * (2.0 * m1.grad).abs().max().item()


@_UNRUN
def test_registrar_with_fused_similarity_follows_the_unfused_trajectory(cuda):
    from tests._scene import make_drr
    from xvr_b200.registrar import Registrar

    res = []
    for fused in (False, True):
        drr = make_drr(96, 64)
        rot0 = torch.tensor([[0.20, -0.10, 0.05]], device=cuda)
        xyz0 = torch.tensor([[5.0, 800.0, -10.0]], device=cuda)
        with torch.no_grad():
            gt = drr(xvr_b200.convert(rot0, xyz0, parameterization="euler_angles", convention="ZXY"))
        init = xvr_b200.convert(rot0 + torch.tensor([[0.06, -0.05, 0.04]], device=cuda),
                                xyz0 + torch.tensor([[8.0, 12.0, -6.0]], device=cuda),
                                parameterization="euler_angles", convention="ZXY")
        pose, info = Registrar(drr, scales="1", n_itrs="60", fused_similarity=fused).run(gt, init)
        res.append((pose.matrix.clone(), info))
    a, b = res[0][1]["nccs"], res[1][1]["nccs"]
    assert len(a) == len(b)
    assert max(abs(u - v) for u, v in zip(a, b)) < 2e-3
    assert b[-1] > b[0]
    assert (res[0][0] - res[1][0]).abs().max().item() < 0.5  # mm / unit rotation entries


# ------------------------------------------------------------------------------------------ staged-brick trilinear
# csrc/trilinear_staged.cu waits on an mbarrier (bounded spin, but a first run belongs under a watchful timeout):
# these tests only run when XVR_B200_RUN_UNVALIDATED=1, e.g. from scripts/gpu_round2_first.sh.
import os  # noqa: E402

_STAGED = pytest.mark.skipif(os.environ.get("XVR_B200_RUN_UNVALIDATED") != "1",
                             reason="staged-brick kernel: never run on a GPU; set XVR_B200_RUN_UNVALIDATED=1")


def _both_kernels(drr, rot, xyz, monkeypatch, stages="1"):
    """(texture kernel, staged kernel with `stages` buffers): image, pose gradients, staging statistics."""
    from xvr_b200 import renderers

    outs = []
    for staged in ("0", stages):
        monkeypatch.setenv("XVR_B200_STAGED", staged)
        stats = torch.zeros(3, dtype=torch.int64, device=rot.device)
        renderers._staged_stats["tensor"] = stats if staged != "0" else None
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
        wimg = torch.rand(img.shape, generator=torch.Generator().manual_seed(4)).to(img.device)
        (img * wimg).sum().backward()
        outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone(), stats.tolist()))
    renderers._staged_stats["tensor"] = None
    return outs


@_STAGED
@pytest.mark.parametrize("stages", ["1", "2"])
@pytest.mark.parametrize("n,h", [(64, 32), (96, 48), (128, 80)])
def test_staged_bricks_render_bit_identical_images_and_gradients(cuda, monkeypatch, n, h, stages):
    from tests._scene import make_drr, pose_params

    drr = make_drr(n, h)
    rot, xyz = pose_params(3, seed=31)
    (img0, gr0, gx0, _), (img1, gr1, gx1, stats) = _both_kernels(drr, rot, xyz, monkeypatch, stages)
    assert torch.equal(img1, img0) and torch.equal(gr1, gr0) and torch.equal(gx1, gx0)
    shared, glob, timeouts = stats
    assert timeouts == 0
    assert shared > 10 * glob  # the bricks serve (nearly) every sample


@_STAGED
@pytest.mark.parametrize("stages", ["1", "2"])
def test_staged_bricks_edge_poses_and_partial_tiles(cuda, monkeypatch, stages):
    """Non-square detector that is not a multiple of the 16 x 16 tile, anisotropic voxels, a volume whose rows are
    not 16-byte multiples (global-memory service), rays missing / grazing the volume, a source inside it."""
    import numpy as np

    from tests._scene import make_drr
    from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
    from xvr_b200.data import read

    drr = make_drr(64, 32)
    a, b = _both_kernels(drr, torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda), monkeypatch,
                         stages)
    assert all(torch.equal(u, v) for u, v in zip(a[:3], b[:3])) and b[3][2] == 0

    from tests._scene import pose_params

    for shape in ((40, 64, 52), (40, 64, 50)):
        vol = torch.rand(*shape, generator=torch.Generator().manual_seed(3)) * 1000 - 500
        drr = xvr_b200.DRR(read(vol, affine=np.diag([2.0, 1.5, 2.5, 1.0])), 1020.0, 24, 6.0, width=40, dely=5.0,
                           x0=7.0, y0=-11.0, renderer="trilinear", reverse_x_axis=True).to(cuda)
        rot, xyz = pose_params(3, seed=5)
        a, b = _both_kernels(drr, rot, xyz, monkeypatch, stages)
        assert all(torch.equal(u, v) for u, v in zip(a[:3], b[:3])) and b[3][2] == 0
        if shape[2] % 4:
            assert b[3][0] == 0  # rows of 50 floats cannot be bulk-copied: everything comes from global memory
