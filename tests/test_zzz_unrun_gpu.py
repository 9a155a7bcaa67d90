"""GPU tests of code written AFTER round 1's GPU budget was spent -- never run on a device yet.

Each is a non-strict ``xfail`` (a pass is reported as XPASS, a failure does not fail the suite) and the file sorts
after every validated test, so that nothing here can mask or disturb them.  Round 2: run, fix what fails, move the
tests next to their validated neighbours and drop the markers.

  * the gather (version 1) dL/dvolume kernel, whose pixel window was rewritten twice after its last full GPU run
    (DESIGN.md 5.6); the brick kernel (the default) passes every one of these cases;
  * the staged-brick trilinear kernel, additionally behind XVR_B200_RUN_UNVALIDATED=1.
(The fused registration similarity started here too; it passed on the B200 and moved to tests/test_regsim_gpu.py.)
"""

import pytest
import torch

import xvr_b200
from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ, _volume_gradient_vs_oracle
from xvr_b200._lib import call

pytestmark = pytest.mark.gpu

_UNRUN = pytest.mark.xfail(strict=False, reason="written after round 1's GPU budget was spent; not yet run on a B200")


@pytest.fixture
def gather_kernel():
    call("xvr_set_volgrad_version", 1)
    yield
    call("xvr_set_volgrad_version", 2)


@_UNRUN
def test_gather_volume_gradient_edge_poses(cuda, gather_kernel):
    """Measured on the B200: 5.8e-2 off with the original pixel window, 3.0e-4 with the projected-corner window near
    the source only; the window is now the projected-corner box everywhere (exact in the CPU emulation)."""
    from tests._scene import make_drr

    drr = make_drr(64, 32)
    _volume_gradient_vs_oracle(drr, torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda))


@_UNRUN
@pytest.mark.parametrize("n,h,b", [(24, 16, 3), (40, 33, 2), (50, 64, 2)])
def test_gather_volume_gradient_cubic_volumes(cuda, gather_kernel, n, h, b):
    from tests._scene import make_drr, pose_params

    _volume_gradient_vs_oracle(make_drr(n, h), *pose_params(b, seed=14))


@_UNRUN
def test_gather_volume_gradient_anisotropic_voxels_offset_reversed_detector(cuda, gather_kernel):
    from tests._scene import pose_params
    from tests.test_zz_full_size_gpu import _anisotropic_drr

    _volume_gradient_vs_oracle(_anisotropic_drr(cuda), *pose_params(3, seed=5))


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


# ------------------------------------------------------------------------------------------ Siddon integer walk
@pytest.fixture
def siddon_walk():
    call("xvr_set_siddon_walk", 1)
    yield
    call("xvr_set_siddon_walk", 0)


@_UNRUN
@pytest.mark.parametrize("n,h,b,shift", [(64, 48, 4, 0.5), (40, 33, 3, 0.0), (256, 128, 2, 0.5)])
def test_siddon_walk_traversal_is_bit_exact(cuda, siddon_walk, n, h, b, shift):
    """xvr_set_siddon_walk(1): indices from the integer walk + one-compare certificate, against the oracle's
    reconstruction of the reference's indices (same assertions as test_siddon_gpu.test_traversal_is_bit_exact)."""
    import oracle
    from tests._scene import make_drr, pose_params
    from tests.test_siddon_gpu import _rays, _trace

    drr = make_drr(n, h, renderer="siddon", voxel_shift=shift)
    rot, xyz = pose_params(b, seed=11)
    source, target, _ = _rays(drr, rot, xyz)
    ref_idx, ref_seg = oracle.siddon_segments(tuple(drr.density.shape), source, target, voxel_shift=shift)
    M = ref_idx.shape[-1]
    idx, seg, cnt = _trace(drr.density, source, target, shift, M + 8)
    ref_valid = torch.diff(oracle.siddon_alphas(source, target, tuple(drr.density.shape), shift, 1e-8), dim=-1)
    ref_cnt = (~ref_valid.isnan()).sum(-1).to(torch.int32)
    assert torch.equal(cnt, ref_cnt)
    live = torch.arange(M, device=cuda)[None, None] < ref_cnt[..., None]
    pos = live & (ref_seg > 0)  # ties between axes may be emitted in either order: zero-length segments are exempt
    assert torch.equal(seg[..., :M][live], ref_seg[live])
    assert torch.equal(idx[..., :M][pos].long(), ref_idx[pos])


@_UNRUN
def test_siddon_walk_renders_bit_identical_images_and_gradients(cuda):
    from tests._scene import make_drr, pose_params

    drr = make_drr(128, 64, renderer="siddon")
    rot, xyz = pose_params(3, seed=3)
    outs = []
    for on in (0, 1):
        call("xvr_set_siddon_walk", on)
        try:
            r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
            img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
            img.sum().backward()
            outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone()))
        finally:
            call("xvr_set_siddon_walk", 0)
    for u, v in zip(*outs):
        assert torch.equal(u, v)
