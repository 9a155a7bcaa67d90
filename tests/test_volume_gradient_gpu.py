"""dL/dvolume beyond the fused trilinear path: the Siddon renderer's brick-local adjoint
(atomics-free, deterministic) and the ray entry points of both renderers (RED.ADD scatter, with and without label
channels), all against autograd through the oracle's grid_sample formulation."""

import numpy as np
import pytest
import torch

import oracle
import xvr_b200
from tests._scene import make_drr, pose_params, rel_l2
from tests.test_siddon_gpu import _rays

pytestmark = pytest.mark.gpu


def _oracle_volume_gradient(drr, rot, xyz, wimg, renderer, mask=None, **kw):
    vref = drr.density.detach().clone().requires_grad_()
    pose = oracle.pose_from_params(rot, xyz, "euler_angles", "ZXY")
    d = drr.detector
    img = oracle.drr_forward(vref, drr._affine_inverse[None], pose, reorient=d._reorient, height=d.height,
                             width=d.width, delx=d.delx, dely=d.dely, x0=d.x0, y0=d.y0, sdd=d.sdd,
                             reverse_x_axis=d.reverse_x_axis, renderer=renderer, mask=mask, **kw)
    (img * wimg).sum().backward()
    return vref.grad


@pytest.mark.parametrize("n,h,b,shift", [(24, 16, 3, 0.5), (40, 33, 2, 0.5), (48, 40, 2, 0.0)])
def test_siddon_fused_volume_gradient_matches_oracle_and_is_deterministic(cuda, n, h, b, shift):
    drr = make_drr(n, h, renderer="siddon", voxel_shift=shift)
    rot, xyz = pose_params(b, seed=14)
    wimg = torch.rand(b, 1, h, h, generator=torch.Generator().manual_seed(2)).to(cuda)
    vol = drr.density.detach().clone().requires_grad_()
    drr.density = vol
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    grads = []
    for _ in range(2):
        vol.grad = None
        (drr(pose) * wimg).sum().backward()
        grads.append(vol.grad.clone())
    assert torch.equal(grads[0], grads[1])  # no atomics: bit-reproducible
    ref = _oracle_volume_gradient(drr, rot, xyz, wimg, "siddon", voxel_shift=shift)
    assert ref.abs().max().item() > 0
    # the fused entry generates its rays from the composed camera -> voxel matrix: a handful of segments at voxel faces
    # resolve to the neighbouring voxel, so the bar is on the norm; the ray entry point below is held to 1e-5
    assert rel_l2(grads[0], ref) < 2e-3
    assert (grads[0].sum() - ref.sum()).abs().item() < 1e-4 * ref.sum().abs().item()  # the chord lengths telescope


def test_siddon_fused_volume_gradient_edge_poses_anisotropic(cuda):
    """Rays missing / grazing the volume, the source inside it, anisotropic voxels, non-square reversed detector."""
    from tests.test_zz_full_size_gpu import EDGE_ROT, EDGE_XYZ
    from xvr_b200.data import read

    drr = make_drr(48, 24, renderer="siddon")
    rot, xyz = torch.tensor(EDGE_ROT, device=cuda), torch.tensor(EDGE_XYZ, device=cuda)
    # the edge translations are tuned for a 64^3 scene; they still cover miss / graze / inside for 48^3
    cases = [(drr, rot, xyz)]
    vol = torch.rand(40, 64, 52, generator=torch.Generator().manual_seed(3)) * 1000 - 500
    aniso = xvr_b200.DRR(read(vol, affine=np.diag([2.0, 1.5, 2.5, 1.0])), 1020.0, 24, 6.0, width=40, dely=5.0, x0=7.0,
                         y0=-11.0, renderer="siddon", reverse_x_axis=True).to(cuda)
    cases.append((aniso, *pose_params(3, seed=5)))
    for d, r, x in cases:
        b = r.shape[0]
        wimg = torch.rand(b, 1, d.detector.height, d.detector.width, generator=torch.Generator().manual_seed(3)).to(cuda)
        v = d.density.detach().clone().requires_grad_()
        d.density = v
        (d(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY")) * wimg).sum().backward()
        ref = _oracle_volume_gradient(d, r, x, wimg, "siddon")
        assert torch.isfinite(v.grad).all()
        assert rel_l2(v.grad, ref) < 5e-3, rel_l2(v.grad, ref)


@pytest.mark.parametrize("renderer", ["trilinear", "siddon"])
@pytest.mark.parametrize("labels", [False, True])
def test_ray_entry_volume_gradient_matches_oracle(cuda, renderer, labels):
    """drr.renderer(volume.requires_grad, source, target, raylen, mask=) -- the call of
    /root/reference/src/xvr/model/trainer.py:288 -- back-propagates to the volume (and to the rays in the same
    backward), per-channel upstream gradients included."""
    n, h, b = 32, 20, 2
    drr = make_drr(n, h, renderer=renderer, with_labels=labels)
    rot, xyz = pose_params(b, seed=9)
    source, target, raylen = _rays(drr, rot, xyz)
    mask = drr.mask if labels else None
    C = int(mask.max()) + 1 if labels else 1
    wimg = torch.rand(b, C, h * h, generator=torch.Generator().manual_seed(6)).to(cuda)

    vol = drr.density.detach().clone().requires_grad_()
    t1 = target.clone().requires_grad_()
    img = drr.renderer(vol, source, t1, raylen, mask=mask)
    assert img.shape == (b, C, h * h)
    (img * wimg).sum().backward()

    vref = drr.density.detach().clone().requires_grad_()
    t2 = target.clone().requires_grad_()
    render = oracle.trilinear_render if renderer == "trilinear" else oracle.siddon_render
    ref = render(vref, source, t2, raylen, mask=mask)
    (ref * wimg).sum().backward()
    assert rel_l2(img.detach(), ref.detach()) < 1e-4
    assert rel_l2(vol.grad, vref.grad) < 1e-5, rel_l2(vol.grad, vref.grad)
    assert rel_l2(t1.grad, t2.grad) < 2e-3


def test_ray_entry_volume_gradient_with_collapsed_channels(cuda):
    """The training loop's img.sum(dim=1) hands back an expanded (stride-0) upstream gradient; with d/dvolume requested
    it is materialised per channel for the re-march."""
    drr = make_drr(32, 16, renderer="trilinear", with_labels=True)
    rot, xyz = pose_params(2, seed=4)
    source, target, raylen = _rays(drr, rot, xyz)
    vol = drr.density.detach().clone().requires_grad_()
    drr.renderer(vol, source, target, raylen, mask=drr.mask).sum(dim=1).square().sum().backward()
    vref = drr.density.detach().clone().requires_grad_()
    oracle.trilinear_render(vref, source, target, raylen, mask=drr.mask).sum(dim=1).square().sum().backward()
    assert rel_l2(vol.grad, vref.grad) < 1e-5
