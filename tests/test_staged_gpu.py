"""The TMA-staged trilinear kernel (csrc/trilinear_staged.cu, XVR_B200_STAGED=1) against the texture kernel: same
arithmetic, same summation order -> bit-identical images and pose gradients; every wait in it is bounded, and the
`stats` counters say how many samples the shared-memory bricks served."""

import numpy as np
import pytest
import torch

import xvr_b200

pytestmark = pytest.mark.gpu


def _both_kernels(drr, rot, xyz, monkeypatch):
    """(texture kernel, staged kernel): image, pose gradients, staging statistics."""
    from xvr_b200 import renderers

    from xvr_b200._lib import options

    outs = []
    for staged in ("0", "1"):
        monkeypatch.setenv("XVR_B200_STAGED", staged)
        stats = torch.zeros(3, dtype=torch.int64, device=rot.device)
        renderers._staged_stats["tensor"] = stats if staged != "0" else None
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        # one lane per ray in the texture kernel too (small launches otherwise split a ray's samples over several
        # lanes, which changes the summation order -- test_sample_slicing_across_lanes_matches_one_lane_per_ray)
        with options(ksplit=0):
            img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
        wimg = torch.rand(img.shape, generator=torch.Generator().manual_seed(4)).to(img.device)
        (img * wimg).sum().backward()
        torch.cuda.synchronize()
        outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone(), stats.tolist()))
    renderers._staged_stats["tensor"] = None
    return outs


@pytest.mark.parametrize("n,h", [(64, 32), (96, 48), (128, 80), (256, 128)])
def test_staged_bricks_render_bit_identical_images_and_gradients(cuda, monkeypatch, n, h):
    from tests._scene import make_drr, pose_params

    drr = make_drr(n, h)
    rot, xyz = pose_params(3, seed=31)
    (img0, gr0, gx0, _), (img1, gr1, gx1, stats) = _both_kernels(drr, rot, xyz, monkeypatch)
    shared, glob, timeouts = stats
    assert timeouts == 0
    assert torch.equal(img1, img0)
    assert torch.equal(gr1, gr0) and torch.equal(gx1, gx0)
    assert shared > 0, stats  # (coarse test detectors put rays many voxels apart: their boxes often exceed a stage)


def test_staged_bricks_no_gradient_variant(cuda, monkeypatch):
    from tests._scene import make_drr, pose_params

    drr = make_drr(96, 64)
    rot, xyz = pose_params(2, seed=7)
    imgs = []
    from xvr_b200._lib import options

    for staged in ("0", "1"):
        monkeypatch.setenv("XVR_B200_STAGED", staged)
        with torch.no_grad(), options(ksplit=0):
            imgs.append(drr(xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")))
    assert torch.equal(imgs[0], imgs[1])


def test_staged_bricks_edge_poses_and_partial_tiles(cuda, monkeypatch):
    """Non-square detector that is not a multiple of the tile, anisotropic voxels, a volume whose rows are not
    16-byte multiples (texture kernel serves it), rays missing / grazing the volume, a source inside it."""
    from tests._scene import make_drr, pose_params
    from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
    from xvr_b200.data import read

    drr = make_drr(64, 32)
    a, b = _both_kernels(drr, torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda), monkeypatch)
    assert b[3][2] == 0
    assert all(torch.equal(u, v) for u, v in zip(a[:3], b[:3]))

    for shape in ((40, 64, 52), (40, 64, 50)):
        vol = torch.rand(*shape, generator=torch.Generator().manual_seed(3)) * 1000 - 500
        drr = xvr_b200.DRR(read(vol, affine=np.diag([2.0, 1.5, 2.5, 1.0])), 1020.0, 24, 6.0, width=40, dely=5.0,
                           x0=7.0, y0=-11.0, renderer="trilinear", reverse_x_axis=True).to(cuda)
        rot, xyz = pose_params(3, seed=5)
        a, b = _both_kernels(drr, rot, xyz, monkeypatch)
        assert b[3][2] == 0
        assert all(torch.equal(u, v) for u, v in zip(a[:3], b[:3]))
        if shape[2] % 4:
            assert b[3][0] == 0  # rows of 50 floats are not TMA-addressable: the texture kernel rendered this


def test_staged_bricks_all_marching_axes_and_directions(cuda, monkeypatch):
    """Views along +-x, +-y, +-z of the volume (every marching axis, forward and backward travel) and oblique ones."""
    from tests._scene import make_drr

    drr = make_drr(96, 48)
    h = 1.5707964
    rot = torch.tensor([[0.0, 0.0, 0.0], [3.1415927, 0.0, 0.0], [h, 0.0, 0.0], [-h, 0.0, 0.0], [0.0, h, 0.0],
                        [0.0, -h, 0.0], [0.7, 0.6, 0.2], [2.3, -0.7, 0.1]], device=cuda)
    xyz = torch.tensor([[0.0, 800.0, 0.0]] * 8, device=cuda)
    a, b = _both_kernels(drr, rot, xyz, monkeypatch)
    assert b[3][2] == 0
    assert all(torch.equal(u, v) for u, v in zip(a[:3], b[:3]))
    assert b[3][0] > 0, b[3]
