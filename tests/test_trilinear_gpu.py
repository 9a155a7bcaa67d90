"""Parity of the CUDA trilinear renderer (through the C-ABI) against the oracle on real grid_sample.

Tolerance: north_star asks for 1e-4 relative fp32 on the DRR; gradients are compared at 2e-3 relative L2
(the oracle's own fp32 autograd through 500-sample sums and atomics is only reproducible to ~1e-3).
"""

import pytest
import torch

import xvr_b200
from tests._scene import make_drr, oracle_render, pose_params, rel_l2

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 2e-3


def _render(drr, rot, xyz, **kw):
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    return drr(pose, **kw)


@pytest.mark.parametrize("n,h,b", [(128, 64, 4), (96, 33, 3)])
def test_forward_matches_oracle(cuda, n, h, b):
    drr = make_drr(n, h)
    rot, xyz = pose_params(b)
    img = _render(drr, rot, xyz)
    ref = oracle_render(drr, rot, xyz)
    assert img.shape == ref.shape == (b, 1, h, h)
    assert rel_l2(img, ref) < FWD_TOL
    assert (img - ref).abs().max().item() < FWD_TOL * ref.abs().max().item()
    assert (img > 0).float().mean() > 0.1


@pytest.mark.parametrize("step", ["span/(n-1)", "span/n", "1/n"])
@pytest.mark.parametrize("n_points", [500, 37])
def test_step_conventions_and_sample_counts(cuda, step, n_points):
    """Every setting of the TRILINEAR_STEP knob and a non-default n_points, fwd + grad."""
    import oracle

    drr = make_drr(48, 24, step=step)
    rot, xyz = pose_params(2, seed=16)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    with torch.no_grad():
        source, target = drr.detector(pose, None)
        raylen = (target - source).norm(dim=-1).unsqueeze(1)
        source, target = drr.affine_inverse(source), drr.affine_inverse(target)
    outs = []
    for fn in (lambda s, t: drr.renderer(drr.density, s, t, raylen, n_points=n_points),
               lambda s, t: oracle.trilinear_render(drr.density, s, t, raylen, n_points=n_points, step=step)):
        s, t = source.clone().requires_grad_(), target.clone().requires_grad_()
        img = fn(s, t)
        img.sum().backward()
        outs.append((img.detach(), s.grad, t.grad))
    assert rel_l2(outs[0][0], outs[1][0]) < FWD_TOL
    assert rel_l2(outs[0][1], outs[1][1]) < GRAD_TOL and rel_l2(outs[0][2], outs[1][2]) < GRAD_TOL


def test_forward_config1_cpu_oracle(cuda):
    """BASELINE config 1: 128^3, 4 poses, 64x64 trilinear DRR, oracle on the CPU (reference's own runnable case)."""
    drr = make_drr(128, 64)
    rot, xyz = pose_params(4, seed=1)
    img = _render(drr, rot, xyz)
    ref = oracle_render(drr.cpu(), rot.cpu(), xyz.cpu())
    assert rel_l2(img.cpu(), ref) < FWD_TOL


def test_forward_nonsquare_anisotropic(cuda):
    import numpy as np

    from xvr_b200.data import read

    g = torch.Generator().manual_seed(3)
    vol = torch.rand(40, 64, 52, generator=g) * 1000 - 500
    affine = np.diag([2.0, 1.5, 2.5, 1.0])
    sub = read(vol, affine=affine)
    drr = xvr_b200.DRR(sub, 1020.0, 24, 6.0, width=40, dely=5.0, x0=7.0, y0=-11.0, renderer="trilinear",
                       reverse_x_axis=True).cuda()
    rot, xyz = pose_params(3, seed=5)
    img = _render(drr, rot, xyz)
    ref = oracle_render(drr, rot, xyz)
    assert img.shape == (3, 1, 24, 40)
    assert rel_l2(img, ref) < FWD_TOL


def test_forward_edge_poses(cuda):
    """Rays missing the volume, grazing it, source inside the volume, axis-aligned rays."""
    drr = make_drr(64, 32)
    rot = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [1.2, 0.3, 0.0], [0.0, 0.0, 0.0], [0.0, 1.5707964, 0.0]],
                       device=cuda)
    xyz = torch.tensor([[0.0, 800.0, 0.0], [400.0, 800.0, 0.0], [0.0, 300.0, 0.0], [10.0, 60.0, -5.0],
                        [128.0, 500.0, 127.5]], device=cuda)
    img = _render(drr, rot, xyz)
    ref = oracle_render(drr, rot, xyz)
    scale = ref.abs().max().item()
    assert (img - ref).abs().max().item() < FWD_TOL * scale
    assert img[1].abs().max().item() == 0.0 or rel_l2(img[1], ref[1]) < 1e-3


def test_labels_to_channels(cuda):
    drr = make_drr(64, 32, with_labels=True)
    rot, xyz = pose_params(3, seed=2)
    img = _render(drr, rot, xyz, mask_to_channels=True)
    ref = oracle_render(drr, rot, xyz, mask=drr.mask)
    assert img.shape == ref.shape and img.shape[1] == int(drr.mask.max()) + 1
    # nearest-label lookups can flip for samples within an ulp of a half-integer: compare per-channel loosely,
    # the channel sum tightly
    assert rel_l2(img.sum(1), ref.sum(1)) < FWD_TOL
    assert rel_l2(img, ref) < 1e-3


@pytest.mark.parametrize("with_labels", [False, True])
def test_pose_gradients_match_oracle(cuda, with_labels):
    drr = make_drr(64, 32, with_labels=with_labels)
    rot, xyz = pose_params(3, seed=4)
    g = torch.Generator().manual_seed(0)
    C = int(drr.mask.max()) + 1 if with_labels else 1
    wimg = torch.rand(3, C, 32, 32, generator=g).to(cuda)

    r1, x1 = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    (_render(drr, r1, x1, mask_to_channels=with_labels) * wimg).sum().backward()
    r2, x2 = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    (oracle_render(drr, r2, x2, mask=drr.mask if with_labels else None) * wimg).sum().backward()
    assert rel_l2(r1.grad, r2.grad) < GRAD_TOL
    assert rel_l2(x1.grad, x2.grad) < GRAD_TOL


def test_ray_gradients_generic_entry(cuda):
    """The drr.renderer(...) call site of trainer.py:288: gradients w.r.t. source, target and ray length."""
    import oracle

    drr = make_drr(64, 32)
    rot, xyz = pose_params(2, seed=6)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    with torch.no_grad():
        source, target = drr.detector(pose, None)
        raylen = (target - source).norm(dim=-1).unsqueeze(1)
        source, target = drr.affine_inverse(source), drr.affine_inverse(target)
    wimg = torch.rand(2, 1, 32 * 32, device=cuda)
    outs = []
    for fn in (lambda s, t, r: drr.renderer(drr.density, s, t, r),
               lambda s, t, r: oracle.trilinear_render(drr.density, s, t, r)):
        s, t, r = (v.clone().requires_grad_() for v in (source, target, raylen))
        (fn(s, t, r) * wimg).sum().backward()
        outs.append((s.grad, t.grad, r.grad))
    for a, b in zip(*outs):
        assert rel_l2(a, b) < GRAD_TOL


def test_jacobian_and_recompute_backward_agree(cuda):
    from xvr_b200._lib import call, ptr, stream

    drr = make_drr(64, 32)
    rot, xyz = pose_params(2, seed=7)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    source, target = drr.detector(pose, None)
    raylen = (target - source).norm(dim=-1).unsqueeze(1).contiguous()
    source, target = drr.affine_inverse(source).contiguous(), drr.affine_inverse(target).contiguous()
    s = source.clone().requires_grad_()
    t = target.clone().requires_grad_()
    gout = torch.rand(2, 1, 1024, device=cuda)
    drr.renderer(drr.density, s, t, raylen).backward(gout)
    B, N = 2, 1024
    gs, gt, gl, work = (torch.empty(B, 1, 3, device=cuda), torch.empty(B, N, 3, device=cuda),
                        torch.empty(B, 1, N, device=cuda), torch.empty(B, 3, N, device=cuda))
    vol = drr.density
    call("xvr_trilinear_rays_bwd", ptr(vol), None, *vol.shape, None, 1, ptr(source), ptr(target), ptr(raylen), B, N, 500,
         0, 1e-8, 32, 32, 3, 4, ptr(gout), ptr(gs), ptr(gt), ptr(gl), ptr(work), None, stream())
    assert rel_l2(gs, s.grad) < 1e-5
    assert rel_l2(gt, t.grad) < 1e-5


def test_fused_and_generic_paths_agree(cuda, monkeypatch):
    """DRR.forward's fused kernel (in-kernel ray generation, dL/dG reduction) against the materialised
    detector -> affine_inverse -> renderer sequence of trainer.py:283-289."""
    drr = make_drr(96, 40, width=56)
    rot, xyz = pose_params(3, seed=12)
    wimg = torch.rand(3, 1, 40, 56, device=cuda)
    res = []
    for fused in ("1", "0"):
        monkeypatch.setenv("XVR_B200_FUSED", fused)
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        img = _render(drr, r, x)
        (img * wimg).sum().backward()
        res.append((img.detach(), r.grad, x.grad))
    assert rel_l2(res[0][0], res[1][0]) < 2e-5
    # the noisy phantom's gradient is a heavily cancelling sum of piecewise-constant voxel differences: moving
    # the ray end points by one fp32 ulp changes it at the 1e-3 level (the smooth-volume test below is tight)
    assert rel_l2(res[0][1], res[1][1]) < 5e-3
    assert rel_l2(res[0][2], res[1][2]) < 5e-3


def _smooth_drr(cuda, n=64, h=32, renderer="trilinear"):
    from xvr_b200.data import read

    ax = torch.linspace(-1, 1, n)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = 800 * torch.exp(-((X - 0.1) ** 2 + (Y + 0.2) ** 2 + Z**2) / 0.18) + 400 * torch.exp(
        -((X + 0.3) ** 2 + (Y - 0.1) ** 2 + (Z - 0.2) ** 2) / 0.08) - 700
    import numpy as np

    sub = read(vol, affine=np.diag([256.0 / n] * 3 + [1.0]))
    return xvr_b200.DRR(sub, 1020.0, h, 1.08821875 * 256 / h, renderer=renderer, reverse_x_axis=False).to(cuda)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_pose_gradients_smooth_volume_tight(cuda, monkeypatch, fused):
    """On a smooth volume the analytic Jacobian must match autograd through grid_sample to 2e-4."""
    monkeypatch.setenv("XVR_B200_FUSED", fused)
    drr = _smooth_drr(cuda)
    rot, xyz = pose_params(3, seed=13)
    wimg = torch.rand(3, 1, 32, 32, device=cuda)
    r1, x1 = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    (_render(drr, r1, x1) * wimg).sum().backward()
    r2, x2 = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    (oracle_render(drr, r2, x2) * wimg).sum().backward()
    assert rel_l2(r1.grad, r2.grad) < 2e-4
    assert rel_l2(x1.grad, x2.grad) < 2e-4


@pytest.fixture
def volgrad_version(request):
    from xvr_b200._lib import options

    with options(volgrad=request.param):
        yield request.param


@pytest.mark.parametrize("volgrad_version", ["brick", "gather"], indirect=True)
@pytest.mark.parametrize("n,h,b", [(24, 16, 3), (40, 33, 2), (50, 64, 2)])
def test_volume_gradient_matches_oracle_and_is_deterministic(cuda, n, h, b, volgrad_version):
    """dL/dvolume (atomics-free: brick-local scatter, and the voxel-centric gather kept as an independent
    cross-check) vs autograd through grid_sample's atomicAdd scatter."""
    import oracle

    drr = make_drr(n, h)
    rot, xyz = pose_params(b, seed=14)
    wimg = torch.rand(b, 1, h, h, device=cuda)
    vol = drr.density.detach().clone().requires_grad_()
    drr.density = vol
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    (drr(pose) * wimg).sum().backward()
    g1 = vol.grad.clone()
    vol.grad = None
    (drr(pose) * wimg).sum().backward()
    assert torch.equal(vol.grad, g1)  # bit-reproducible

    vref = vol.detach().clone().requires_grad_()
    d = drr.detector
    img = oracle.drr_forward(vref, drr._affine_inverse[None], pose.matrix, reorient=d._reorient, height=d.height,
                             width=d.width, delx=d.delx, dely=d.dely, x0=d.x0, y0=d.y0, sdd=d.sdd,
                             reverse_x_axis=d.reverse_x_axis)
    (img * wimg).sum().backward()
    assert rel_l2(g1, vref.grad) < 1e-4
    assert (g1 - vref.grad).abs().max().item() < 1e-4 * vref.grad.abs().max().item()


def test_texture_and_linear_gathers_agree_bitwise(cuda, monkeypatch):
    """The TLD4 path fetches the same fp32 texels as the scalar-load path: images and gradients are identical.  (One lane
    per ray: with several, the lanes slice the TRIMMED sample range, which only the texture handle knows -- same sums,
    grouped differently.)"""
    from xvr_b200._lib import options

    drr = make_drr(64, 32)
    rot, xyz = pose_params(3, seed=9)
    res = []
    for mode in ("tex", "ldg"):
        monkeypatch.setenv("XVR_B200_GATHER", mode)
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        with options(ksplit=0):
            img = _render(drr, r, x)
            img.sum().backward()
        res.append((img.detach(), r.grad, x.grad))
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert torch.equal(a, b)


def test_texture_tracks_volume_updates(cuda):
    """A new or in-place-modified volume tensor must be re-uploaded to the texture (trainer.py:196-197)."""
    drr = make_drr(48, 24)
    rot, xyz = pose_params(2, seed=10)
    img0 = _render(drr, rot, xyz)
    drr.density.mul_(2.0)
    img1 = _render(drr, rot, xyz)
    assert rel_l2(img1, 2 * img0) < 1e-5
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    source, target = drr.detector(pose, None)
    raylen = (target - source).norm(dim=-1).unsqueeze(1)
    source, target = drr.affine_inverse(source), drr.affine_inverse(target)
    for scale in (3.0, 5.0):
        tmp = drr.density * scale  # fresh tensor every step, possibly at a recycled address
        img = drr.renderer(tmp, source, target, raylen).view_as(img1)
        assert rel_l2(img, scale * img1) < 1e-5
        del tmp


def test_sample_slicing_across_lanes_matches_one_lane_per_ray(cuda):
    """Small batches split each ray's samples over 2/4/8 lanes; images and Jacobians agree to summation order."""
    from xvr_b200._lib import options

    drr = make_drr(64, 40, width=24)
    rot, xyz = pose_params(2, seed=15)
    res = []
    for ks in (0, 1, 2, 3, None):
        with options(ksplit=ks):
            r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
            img = _render(drr, r, x)
            img.sum().backward()
            res.append((img.detach(), r.grad, x.grad))
    for img, gr, gx in res[1:]:
        assert rel_l2(img, res[0][0]) < 1e-6
        assert rel_l2(gr, res[0][1]) < 1e-4 and rel_l2(gx, res[0][2]) < 1e-4
    assert torch.equal(res[4][0], res[3][0])  # automatic mode picks 8 lanes per ray for this tiny launch


def test_tile_shapes_give_identical_images(cuda, monkeypatch):
    from xvr_b200._lib import options

    drr = make_drr(64, 64)
    rot, xyz = pose_params(2, seed=8)
    imgs = []
    with options(ksplit=0):  # one lane per ray: the per-ray summation order is then independent of the tile shape
        for tile in ("3,4", "5,5", "0,0", "2,3", "5,8"):
            monkeypatch.setenv("XVR_B200_TILE", tile)
            imgs.append(_render(drr, rot, xyz))
    for im in imgs[1:]:
        assert torch.equal(im, imgs[0])


def test_errors_are_loud(cuda):
    drr = make_drr(32, 16)
    rot, xyz = pose_params(1)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    source, target = drr.detector(pose, None)
    with pytest.raises(xvr_b200._lib.XvrB200Error):  # no CPU path
        drr.renderer(drr.density.cpu(), source.cpu(), target.cpu(), torch.ones(1, 1, 256))
    with pytest.raises(ValueError):
        drr.renderer(drr.density, source, target, torch.ones(1, 1, 3, device=cuda))


@pytest.mark.parametrize("renderer,fused", [("trilinear", "0"), ("trilinear", "1"), ("siddon", "1")])
def test_channel_collapsed_gradient_uses_the_saved_jacobian(cuda, monkeypatch, renderer, fused):
    """trainer.py:292-294 collapses the label channels right after rendering (img.sum(dim=1)); autograd then hands
    the renderer one shared upstream gradient (stride 0 along the channels) and the backward is the Jacobian
    epilogue.  It must equal the per-channel recompute backward fed the same (materialised) gradient, and the
    gradient of the unlabelled render.  On the ray entry points (XVR_B200_FUSED=0, and Siddon's label channels always)
    both backwards see the same rays: 1e-4.  The fused trilinear label render generates its rays in the kernel and
    hands a per-channel gradient to the ray entry point's recompute backward: different rounding of the ray end points
    -> the noisy-phantom gradient bar there."""
    monkeypatch.setenv("XVR_B200_FUSED", fused)
    drr = make_drr(64, 32, renderer=renderer, with_labels=True)
    rot, xyz = pose_params(3, seed=6)
    w = torch.rand(3, 1, 32, 32, generator=torch.Generator().manual_seed(1)).to(cuda)
    grads = []
    for mode in ("collapsed", "materialised", "unlabelled"):
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        pose = xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY")
        if mode == "unlabelled":
            img = drr(pose)
        else:
            img = drr(pose, mask_to_channels=True)
            assert img.shape[1] > 1
            img = img.sum(dim=1, keepdim=True) if mode == "collapsed" else (img * torch.ones_like(img)).sum(1, keepdim=True)
        (img * w).sum().backward()
        grads.append(torch.cat([r.grad, x.grad], 1))
    assert rel_l2(grads[0], grads[1]) < (2e-3 if (renderer, fused) == ("trilinear", "1") else 1e-4)
    # the unlabelled render takes the fused path (rays generated in registers, not read from (B,N,3) tensors): same
    # mathematics, different rounding of the ray end points -> the noisy-phantom gradient bar
    assert rel_l2(grads[0], grads[2]) < 2e-3


# ------------------------------------------------------------------------------------------ empty-space trimming
@pytest.mark.parametrize("ksplit", [0, 2], ids=["1lane", "4lanes"])
@pytest.mark.parametrize("scene", ["phantom", "blob", "dense", "zeros", "labels", "edge"])
def test_empty_space_trimming_is_bit_identical(cuda, scene, ksplit):
    """The forward kernels skip samples outside the box of the volume's non-zero voxels (exact zeros for every sum):
    images, label channels and pose gradients equal the full march bit for bit -- air margins, a small blob in a sea of
    zeros, a volume without a single zero, an all-zero volume, label channels, rays missing / grazing / starting inside."""
    import ctypes

    from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
    from xvr_b200._lib import call, options

    drr = make_drr(64, 40, with_labels=scene == "labels")
    rot, xyz = pose_params(3, seed=17)
    if scene == "blob":
        vol = torch.zeros_like(drr.density)
        vol[20:27, 40:44, 9:30] = torch.rand(7, 4, 21, device=cuda) + 0.1
        drr.density = vol
    elif scene == "dense":
        drr.density = torch.rand_like(drr.density) + 0.5
    elif scene == "zeros":
        drr.density = torch.zeros_like(drr.density)
    elif scene == "edge":
        rot, xyz = torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda)
    outs = []
    for trim in (True, False):
        # several lanes per ray: lane p adds up the samples k = p (mod lanes) -- absolute residue classes, the same
        # samples in the same order whatever the trimming cut away
        with options(trim=trim, ksplit=ksplit):
            r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
            img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"),
                      mask_to_channels=scene == "labels")
            w = torch.rand(img.shape, generator=torch.Generator().manual_seed(1)).to(cuda)
            (img * w).sum().backward()
            outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone()))
    for u, v in zip(*outs):
        assert torch.equal(u, v)
    # the handle reports the box it trims to
    bbox = (ctypes.c_int * 6)()
    call("xvr_volume_bbox", drr.renderer._texture.handle, bbox, None)
    nz = (drr.density != 0).nonzero()
    if scene == "zeros":
        assert list(bbox) == [64, 64, 64, -1, -1, -1]
    else:
        assert list(bbox) == nz.min(0).values.tolist() + nz.max(0).values.tolist()
    if scene == "blob":
        assert outs[0][0].abs().sum() > 0


def test_label_brick_table_gives_the_same_channels(cuda, monkeypatch):
    """XVR_OPT_LABEL_BRICKS answers the nearest-label lookup of uniform bricks from a table: channels and pose gradients
    equal the exact lookup bit for bit (odd volume shape, rays leaving through every face, a source inside the volume)."""
    import numpy as np

    from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
    from xvr_b200.data import read

    g = torch.Generator().manual_seed(3)
    hu = torch.rand(50, 45, 61, generator=g) * 1500 - 700
    lab = torch.zeros(50, 45, 61)
    lab[4:46, 3:40, 2:] = 1
    lab[10:20, 8:30, 5:25] = 2
    lab[30:33, 20:22, 40:61] = 5
    for drr, rot, xyz in (
            (xvr_b200.DRR(read(hu, lab, affine=np.diag([4.0, 5.0, 3.5, 1.0])), 1020.0, 40, 6.0, width=28,
                          renderer="trilinear", reverse_x_axis=False).to(cuda), *pose_params(3, seed=31)),
            (make_drr(64, 40, with_labels=True), torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda))):
        outs = []
        for bricks in ("1", "0"):
            monkeypatch.setenv("XVR_B200_LABEL_BRICKS", bricks)
            r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
            img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"), mask_to_channels=True)
            assert img.shape[1] > 1
            w = torch.rand(img.shape, generator=torch.Generator().manual_seed(1)).to(cuda)
            (img * w).sum().backward()
            outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone()))
        for u, v in zip(*outs):
            assert torch.equal(u, v)
        assert outs[0][0][:, 1:].abs().sum() > 0


def test_fused_label_channels_match_the_ray_entry_point(cuda, monkeypatch):
    """drr(pose, mask_to_channels=True) through xvr_trilinear_drr_fwd_labels (rays generated in the kernel) against the
    materialised-ray entry the reference's Trainer.render_samples sequence uses: channels to 1e-5 (the composed
    camera -> voxel matrix rounds differently from the two-step transform), pose gradients to the noisy-phantom bar --
    for the collapsed channels (the saved Jacobian of the channel sum) and for per-channel upstream gradients (the
    fallback through the ray entry point's recompute backward)."""
    drr = make_drr(64, 40, with_labels=True, width=28)
    rot, xyz = pose_params(3, seed=12)
    for collapsed in (True, False):
        outs = []
        for fused in ("1", "0"):
            monkeypatch.setenv("XVR_B200_FUSED", fused)
            r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
            img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"), mask_to_channels=True)
            assert img.shape[1] > 1
            if collapsed:
                tot = img.sum(dim=1, keepdim=True)
                w = torch.rand(tot.shape, generator=torch.Generator().manual_seed(2)).to(cuda)
                (tot * w).sum().backward()
            else:
                w = torch.rand(img.shape, generator=torch.Generator().manual_seed(2)).to(cuda)
                (img * w).sum().backward()
            outs.append((img.detach().clone(), r.grad.clone(), x.grad.clone()))
        assert rel_l2(outs[0][0], outs[1][0]) < 2e-5
        assert rel_l2(outs[0][1], outs[1][1]) < 2e-3 and rel_l2(outs[0][2], outs[1][2]) < 2e-3
