"""``DRR`` and ``Detector``: drop-in for ``diffdrr.drr.DRR`` / ``diffdrr.detector.Detector``.

Constructor order and attribute surface pinned by xvr: ``DRR(subject, sdd, height, delx, width, dely, x0, y0,
reverse_x_axis=, renderer=, **kw)`` (/root/reference/src/xvr/renderer/load.py:29-41), the direct sub-object
calls ``drr.detector(pose, None)``, ``drr.affine_inverse(points)``, ``drr.renderer(...)``,
``drr.reshape_transform(img, batch_size=)`` (/root/reference/src/xvr/model/trainer.py:283-289), the mutations at
/root/reference/src/xvr/model/utils.py:162-171 (``drr.density = None``, ``register_buffer``, ``.cuda()``),
``set_intrinsics_`` / ``rescale_detector_`` (/root/reference/src/xvr/registrar/base.py:141-157,212) and
``perspective_projection`` / ``inverse_projection`` (/root/reference/src/xvr/metrics/evaluator.py:17-29).
"""

import ctypes
import os

import numpy as np
import torch

from . import _conventions as conv
from .pose import RigidTransform, convert
from .renderers import Siddon, Trilinear

from ._lib import call, cuda_f32, ptr, stream

__all__ = ["DRR", "Detector"]


class _EulerCamera(torch.autograd.Function):
    """(rot, xyz) Euler pose -> (cam2vox, cam2world) (B,3,4) in ONE launch each way: the arithmetic of
    ``convert(rot, xyz, "euler_angles", convention)`` -> ``reorient.compose(pose)`` -> ``affine_inverse @ ...``
    (some 25 small launches, and twice as many in its autograd backward) as csrc/regstep.cu.  Used by
    ``DRR.forward(rot, xyz, parameterization="euler_angles", ...)``, i.e. by ``Registration.forward``: at B = 1 the
    launches around the renderer cost more than the renderer."""

    @staticmethod
    def forward(ctx, rot, xyz, axes, reorient16, affinv16):
        rot, xyz = cuda_f32(rot, "rotation"), cuda_f32(xyz, "translation")
        B = rot.shape[0]
        if rot.shape != (B, 3) or xyz.shape != (B, 3):
            raise ValueError(f"expected (B,3) Euler angles and (B,3) translations; got {tuple(rot.shape)}, {tuple(xyz.shape)}")
        cam2world = torch.empty(B, 3, 4, device=rot.device, dtype=torch.float32)
        cam2vox = torch.empty(B, 3, 4, device=rot.device, dtype=torch.float32)
        ctx.consts = ((ctypes.c_int * 3)(*axes), int(conv.CONVERT_TRANSLATION_IN_ROTATED_FRAME),
                      (ctypes.c_float * 16)(*reorient16), (ctypes.c_float * 16)(*affinv16))
        if B > 0:
            call("xvr_euler_camera_fwd", ptr(rot), ptr(xyz), B, *ctx.consts, ptr(cam2world), ptr(cam2vox), stream())
        ctx.save_for_backward(rot, xyz)
        ctx.mark_non_differentiable(cam2world)  # only feeds the ray length, which a rigid motion leaves unchanged
        return cam2vox, cam2world

    @staticmethod
    def backward(ctx, g_cam2vox, _g_cam2world):
        rot, xyz = ctx.saved_tensors
        B = rot.shape[0]
        grot, gxyz = torch.empty_like(rot), torch.empty_like(xyz)
        if B == 0:
            return grot, gxyz, None, None, None
        call("xvr_euler_camera_bwd", ptr(rot), ptr(xyz), B, *ctx.consts, ptr(cuda_f32(g_cam2vox, "grad")), ptr(grot),
             ptr(gxyz), stream())
        return grot, gxyz, None, None, None


class Detector(torch.nn.Module):
    """C-arm geometry: X-ray source at the camera origin, detector plane at z = sdd.

    ``forward(pose, calibration=None) -> (source (B,1,3), target (B,H*W,3))`` in world mm, rays ordered
    row-major over the (H, W) detector.
    """

    def __init__(self, sdd, height, width, delx, dely, x0, y0, reorient, reverse_x_axis=False, n_subsample=None):
        super().__init__()
        if n_subsample is not None:
            raise NotImplementedError("random ray subsampling is not part of the xvr hot path")
        self.sdd = float(sdd)
        self.height = int(height)
        self.width = int(width)
        self.delx = float(delx)
        self.dely = float(dely)
        self.x0 = float(x0)
        self.y0 = float(y0)
        self.reverse_x_axis = bool(reverse_x_axis)
        self.register_buffer("_reorient", torch.as_tensor(reorient, dtype=torch.float32))
        self.register_buffer("source", torch.zeros(1, 1, 3))
        self.register_buffer("target", self._make_target())

    @property
    def reorient(self):
        return RigidTransform(self._reorient)

    def pixel_basis(self):
        """Camera-frame detector point of pixel (i, j) = origin + i*row_step + j*col_step (3-vectors)."""
        h_off = 1.0 if self.height % 2 else 0.5
        w_off = 1.0 if self.width % 2 else 0.5
        sign_s = conv.DET_SIGN_S * (-1.0 if self.reverse_x_axis else 1.0)
        sign_t = conv.DET_SIGN_T
        s0 = sign_s * ((-self.width) // 2 + w_off)
        t0 = sign_t * ((-self.height) // 2 + h_off)
        origin = (s0 * self.delx + self.x0, t0 * self.dely + self.y0, self.sdd)
        row_step = (0.0, sign_t * self.dely, 0.0)
        col_step = (sign_s * self.delx, 0.0, 0.0)
        return origin, row_step, col_step

    def _make_target(self):
        h_off = 1.0 if self.height % 2 else 0.5
        w_off = 1.0 if self.width % 2 else 0.5
        t = torch.arange(-self.height // 2, self.height // 2, dtype=torch.float32) + h_off
        s = torch.arange(-self.width // 2, self.width // 2, dtype=torch.float32) + w_off
        t = conv.DET_SIGN_T * t
        s = conv.DET_SIGN_S * s
        if self.reverse_x_axis:
            s = -s
        tt, ss = torch.meshgrid(t, s, indexing="ij")
        target = torch.stack([ss * self.delx + self.x0, tt * self.dely + self.y0, torch.full_like(ss, self.sdd)], -1)
        return target.reshape(1, -1, 3)

    @property
    def calibration(self):
        """Pixel-grid -> camera-frame scaling (already folded into ``target``)."""
        return RigidTransform(torch.tensor(
            [[self.delx, 0, 0, self.x0], [0, self.dely, 0, self.y0], [0, 0, self.sdd, 0], [0, 0, 0, 1.0]],
            device=self._reorient.device))

    @property
    def intrinsic(self):
        """3x3 pinhole intrinsic matrix in pixel units."""
        return torch.tensor(
            [[self.sdd / self.delx, 0.0, self.x0 / self.delx + self.width / 2],
             [0.0, self.sdd / self.dely, self.y0 / self.dely + self.height / 2],
             [0.0, 0.0, 1.0]], device=self._reorient.device)

    def forward(self, extrinsic, calibration=None):
        if calibration is not None:
            raise NotImplementedError("a custom calibration transform is not part of the xvr hot path")
        pose = self.reorient.compose(extrinsic)
        B = len(pose)
        source = pose(self.source.expand(B, -1, -1))
        target = pose(self.target.expand(B, -1, -1))
        return source, target


class DRR(torch.nn.Module):
    """Differentiable X-ray renderer: ``drr(pose) -> (B, C, H, W)``."""

    def __init__(self, subject, sdd, height, delx, width=None, dely=None, x0=0.0, y0=0.0, p_subsample=None,
                 reshape=True, reverse_x_axis=True, patch_size=None, renderer="siddon", persistent=True,
                 voxel_shift=conv.SIDDON_VOXEL_SHIFT_DEFAULT, **renderer_kwargs):
        super().__init__()
        if p_subsample is not None or patch_size is not None:
            raise NotImplementedError("p_subsample / patch_size are not needed: the fused kernels hold no "
                                      "(B,N,n_points) intermediates")
        width = height if width is None else width
        dely = delx if dely is None else dely
        self.subject = subject
        self.reshape = reshape
        reorient = subject.reorient if subject.reorient is not None else torch.eye(4)
        self.detector = Detector(sdd, height, width, delx, dely, x0, y0, reorient, reverse_x_axis=reverse_x_axis)

        affine = torch.as_tensor(np.asarray(subject.volume.affine), dtype=torch.float32)
        self.register_buffer("_affine", affine, persistent=persistent)
        self.register_buffer("_affine_inverse", torch.linalg.inv(affine.double()).float(), persistent=persistent)
        density = subject.density if subject.density is not None else subject.volume.data[0]
        self.register_buffer("density", density.to(torch.float32).squeeze(), persistent=persistent)
        if subject.mask is not None:
            self.register_buffer("mask", subject.mask.data[0].to(torch.uint8), persistent=persistent)

        if renderer == "siddon":
            self.renderer = Siddon(voxel_shift=voxel_shift, **renderer_kwargs)
        elif renderer == "trilinear":
            self.renderer = Trilinear(**renderer_kwargs)
        else:
            raise ValueError(f"renderer must be 'siddon' or 'trilinear', not {renderer!r}")
        self.renderer.detector_hw = (self.detector.height, self.detector.width)

    # ------------------------------------------------------------------ geometry helpers
    @property
    def affine(self):
        return RigidTransform(self._affine)

    @property
    def affine_inverse(self):
        return RigidTransform(self._affine_inverse)

    @property
    def device(self):
        return self._affine.device

    def reshape_transform(self, img, batch_size):
        if self.reshape:
            # (B, C, H*W) -> (B, C, H, W); the channel count is spelled out so that an empty batch reshapes too
            return img.view(batch_size, img.shape[1], self.detector.height, self.detector.width)
        return img

    # ------------------------------------------------------------------ rendering
    def forward(self, *args, parameterization=None, convention=None, calibration=None, mask_to_channels=False,
                **kwargs):
        """Render at ``pose`` (a RigidTransform) or at ``(rot, xyz, parameterization=, convention=)``."""
        if self.density is None:
            raise RuntimeError("drr.density was unloaded; call drr.renderer(volume, ...) directly "
                               "(as xvr's Trainer.render_samples does) or restore it")
        mask = getattr(self, "mask", None) if mask_to_channels else None
        # label channels ride the fused path of the trilinear renderer (no volume gradient there: the ray entry point
        # has the scatter for it)
        fused_mask_ok = mask is None or (isinstance(self.renderer, Trilinear) and not self.density.requires_grad)
        fused = (fused_mask_ok and calibration is None and isinstance(self.renderer, (Trilinear, Siddon))
                 and self.reshape and os.environ.get("XVR_B200_FUSED", "1") == "1")
        if mask is not None and fused:
            kwargs = dict(kwargs, mask=mask)
        degrees = kwargs.pop("degrees", False)
        if (fused and parameterization == "euler_angles" and len(args) == 2 and args[0].is_cuda and args[0].dim() == 2
                and conv.COMPOSE_APPLIES_SELF_FIRST and os.environ.get("XVR_B200_FUSED_POSE", "1") == "1"):
            # Euler parameters straight to camera matrices (one launch), then the fused renderer
            rot, xyz = args
            axes = ["XYZ".index(c) for c in convention]
            reorient16, affinv16 = self._host_matrices()
            cam2vox, cam2world = _EulerCamera.apply(torch.deg2rad(rot) if degrees else rot, xyz, axes, reorient16,
                                                    affinv16)
            img = self.renderer.render_drr(self.density, cam2vox, cam2world, self.detector, **kwargs)
            return self.reshape_transform(img, batch_size=rot.shape[0])
        if (fused and parameterization is not None and len(args) == 2 and conv.COMPOSE_APPLIES_SELF_FIRST):
            from .pose import POSE_KERNEL_KINDS, _kernel_eligible, _PoseKernel  # noqa: PLC0415

            if _kernel_eligible(args[0], args[1], parameterization):
                # any closed-form parameterisation straight to camera matrices (one launch), then the fused renderer
                rot, xyz = args
                cam2vox, cam2world = _PoseKernel.apply(rot, xyz, POSE_KERNEL_KINDS[parameterization], convention,
                                                       degrees, self._host_matrices())
                img = self.renderer.render_drr(self.density, cam2vox, cam2world, self.detector, **kwargs)
                return self.reshape_transform(img, batch_size=rot.shape[0])
        if parameterization is None:
            (pose,) = args
        else:
            pose = convert(*args, parameterization=parameterization, convention=convention, degrees=degrees)
        if fused:
            # fused path: one kernel generates the rays, marches them and (if needed) emits the pose Jacobian
            cam2world = self.detector.reorient.compose(pose).matrix
            cam2vox = self._affine_inverse @ cam2world
            img = self.renderer.render_drr(self.density, cam2vox[:, :3].contiguous(), cam2world[:, :3].contiguous(),
                                           self.detector, **kwargs)
            return self.reshape_transform(img, batch_size=len(pose))
        source, target = self.detector(pose, calibration)
        raylen = (target - source).norm(dim=-1).unsqueeze(1)
        source = self.affine_inverse(source)
        target = self.affine_inverse(target)
        img = self.renderer(self.density, source, target, raylen, mask=mask, **kwargs)
        return self.reshape_transform(img, batch_size=len(pose))

    def _host_matrices(self):
        """Host copies (tuples of 16 floats) of the reorientation and the affine inverse for the one-launch pose
        kernels; refreshed when either buffer is replaced or modified.  The device->host read happens on the first
        call (before any CUDA-graph capture: captures are preceded by a warm-up run)."""
        r, a = self.detector._reorient, self._affine_inverse
        key = (r.data_ptr(), r._version, a.data_ptr(), a._version)
        cache = getattr(self, "_host_matrix_cache", None)
        if cache is None or cache[0] != key:
            cache = (key, tuple(r.detach().double().cpu().flatten().tolist()),
                     tuple(a.detach().double().cpu().flatten().tolist()))
            self._host_matrix_cache = cache
        return cache[1], cache[2]

    # ------------------------------------------------------------------ intrinsics
    def set_intrinsics_(self, sdd=None, delx=None, dely=None, x0=None, y0=None, height=None, width=None):
        d = self.detector
        new = Detector(
            d.sdd if sdd is None else sdd,
            d.height if height is None else height,
            d.width if width is None else width,
            d.delx if delx is None else delx,
            d.dely if dely is None else dely,
            d.x0 if x0 is None else x0,
            d.y0 if y0 is None else y0,
            d._reorient,
            reverse_x_axis=d.reverse_x_axis,
        ).to(d._reorient.device)
        self.detector = new
        self.renderer.detector_hw = (new.height, new.width)

    def rescale_detector_(self, scale):
        d = self.detector
        self.set_intrinsics_(height=int(d.height * scale), width=int(d.width * scale), delx=d.delx / scale,
                             dely=d.dely / scale)

    # ------------------------------------------------------------------ projections
    def perspective_projection(self, pose, pts):
        """World points (B,N,3) -> detector pixel coordinates (B,N,2) for the camera at ``pose``."""
        extrinsic = self.detector.reorient.compose(pose).inverse()
        x = extrinsic(pts)
        x = torch.einsum("ij,bnj->bni", self.detector.intrinsic.to(x), x)
        z = x[..., -1:].clone()
        x = x / z
        if self.detector.reverse_x_axis:
            x[..., 0] = self.detector.width - x[..., 0]
        return x[..., :2].flip(-1)

    def inverse_projection(self, pose, pts):
        """Detector pixel coordinates (B,N,2) -> world points on the detector plane (B,N,3)."""
        extrinsic = self.detector.reorient.compose(pose)
        pts = pts.flip(-1)
        if self.detector.reverse_x_axis:
            pts = pts.clone()
            pts[..., 0] = self.detector.width - pts[..., 0]
        x = self.detector.sdd * torch.einsum(
            "ij,bnj->bni", self.detector.intrinsic.inverse().to(pts),
            torch.cat([pts, torch.ones_like(pts[..., :1])], -1))
        return extrinsic(x)
