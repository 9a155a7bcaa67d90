"""Ray-march renderers: drop-in for ``diffdrr.renderers.{Trilinear, Siddon}`` on hand-written sm_100a kernels.

Call signature pinned by /root/reference/src/xvr/model/trainer.py:288:
``renderer(volume (D0,D1,D2), source (B,1,3), target (B,N,3), raylen (B,1,N), mask=labelmap|None) -> (B,C,N)``
with ``source``/``target`` in voxel-index coordinates.  The modules are autograd-transparent: gradients flow
to ``source``, ``target``, ``raylen`` (pose path) and, on request, to ``volume``.
"""

import os

import torch

from . import _conventions as conv
from . import _lib
from ._lib import call, cuda_f32, ptr, stream

__all__ = ["Trilinear", "Siddon"]


def _tile_shape():
    """(lane_w_log2, cta_w_log2): 8x4-pixel warps in 16x16-pixel CTAs unless overridden for tuning."""
    env = os.environ.get("XVR_B200_TILE")
    if env:
        lw, cw = (int(v) for v in env.split(","))
        return lw, cw
    return 3, 4


class _LabelCache:
    """uint8 copy of a label volume and its channel count, refreshed when the source tensor changes."""

    def __init__(self):
        self.key = None
        self.labels = None
        self.channels = 1

    def get(self, mask):
        key = (mask.data_ptr(), mask._version, tuple(mask.shape), mask.dtype, mask.device)
        if key != self.key:
            hi = int(mask.max().item())  # same host sync as the reference's `int(mask.max()) + 1`
            lo = int(mask.min().item())
            if lo < 0 or hi > 254:
                raise _lib.XvrB200Error(f"label volume must hold integers in [0, 254]; got [{lo}, {hi}]")
            self.labels = mask.to(torch.uint8).contiguous()
            self.channels = hi + 1
            self.key = key
        return self.labels, self.channels


def _check_rays(volume, source, target, raylen):
    if volume.dim() != 3:
        raise ValueError(f"volume must be (D0,D1,D2); got {tuple(volume.shape)}")
    if target.dim() != 3 or target.shape[-1] != 3:
        raise ValueError(f"target must be (B,N,3); got {tuple(target.shape)}")
    B, N, _ = target.shape
    if source.shape != (B, 1, 3):
        raise ValueError(f"source must be (B,1,3) = {(B, 1, 3)}; got {tuple(source.shape)}")
    if raylen.shape != (B, 1, N):
        raise ValueError(f"ray length must be (B,1,N) = {(B, 1, N)}; got {tuple(raylen.shape)}")
    return B, N


class _TrilinearRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, volume, source, target, raylen, labels, C, n_points, step_mode, eps, det_hw):
        volume, source, target, raylen = (
            cuda_f32(volume, "volume"), cuda_f32(source, "source"), cuda_f32(target, "target"),
            cuda_f32(raylen, "raylen"))
        B, N = _check_rays(volume, source, target, raylen)
        D0, D1, D2 = volume.shape
        lw, cw = _tile_shape()
        det_h, det_w = det_hw if det_hw is not None and det_hw[0] * det_hw[1] == N else (0, 0)
        need_pose_grad = any(ctx.needs_input_grad[1:4])
        out = torch.empty(B, C, N, device=volume.device, dtype=torch.float32)
        jac = None
        if need_pose_grad and labels is None:
            jac = torch.empty(B, 7, N, device=volume.device, dtype=torch.float32)
        call("xvr_trilinear_rays_fwd", ptr(volume), D0, D1, D2, ptr(labels), C, ptr(source), ptr(target),
             ptr(raylen), B, N, n_points, step_mode, eps, det_h, det_w, lw, cw, ptr(out), ptr(jac), stream())
        ctx.cfg = (C, n_points, step_mode, eps, det_h, det_w, lw, cw)
        if jac is not None:
            ctx.save_for_backward(jac)
            ctx.mode = "jac"
        elif need_pose_grad:
            ctx.save_for_backward(volume, source, target, raylen, labels)
            ctx.mode = "recompute"
        if ctx.needs_input_grad[0]:
            raise _lib.XvrB200Error("d/dvolume of the trilinear renderer is not available yet")
        return out

    @staticmethod
    def backward(ctx, gout):
        C, n_points, step_mode, eps, det_h, det_w, lw, cw = ctx.cfg
        gout = cuda_f32(gout, "grad_output")
        B, _, N = gout.shape
        dev = gout.device
        gsource = torch.empty(B, 1, 3, device=dev, dtype=torch.float32)
        gtarget = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        graylen = torch.empty(B, 1, N, device=dev, dtype=torch.float32)
        work = torch.empty(B, 3, N, device=dev, dtype=torch.float32)
        gvol = None
        if ctx.mode == "jac":
            (jac,) = ctx.saved_tensors
            call("xvr_rays_jac_bwd", ptr(jac), ptr(gout), B, N, ptr(gsource), ptr(gtarget), ptr(graylen),
                 ptr(work), stream())
        else:
            volume, source, target, raylen, labels = ctx.saved_tensors
            D0, D1, D2 = volume.shape
            call("xvr_trilinear_rays_bwd", ptr(volume), D0, D1, D2, ptr(labels), C, ptr(source), ptr(target),
                 ptr(raylen), B, N, n_points, step_mode, eps, det_h, det_w, lw, cw, ptr(gout), ptr(gsource),
                 ptr(gtarget), ptr(graylen), ptr(work), stream())
        return gvol, gsource, gtarget, graylen, None, None, None, None, None, None


class Trilinear(torch.nn.Module):
    """Trilinear ray-marching renderer (``diffdrr.renderers.Trilinear``), n_points samples per ray."""

    def __init__(self, near=0.0, far=1.0, mode="bilinear", filter_intersections_outside_volume=True,
                 eps=conv.RENDER_EPS, step=conv.TRILINEAR_STEP):
        super().__init__()
        if near != 0.0 or far != 1.0 or mode != "bilinear" or not filter_intersections_outside_volume:
            raise NotImplementedError(
                "xvr_b200.Trilinear implements the configuration xvr uses: near=0, far=1, mode='bilinear', "
                "filter_intersections_outside_volume=True")
        self.eps = eps
        self.step = step
        self.detector_hw = None  # set by DRR so that warps map to compact detector tiles
        self._labels = _LabelCache()

    def dims(self, volume):
        return torch.tensor(volume.shape).to(volume) - 1

    def forward(self, volume, source, target, img, n_points=conv.TRILINEAR_N_POINTS, align_corners=True,
                mask=None):
        if not align_corners:
            raise NotImplementedError("xvr_b200.Trilinear implements align_corners=True (the DiffDRR default)")
        labels, C = (None, 1) if mask is None else self._labels.get(mask)
        return _TrilinearRays.apply(volume, source, target, img, labels, C, int(n_points),
                                    conv.STEP_MODES[self.step], float(self.eps), self.detector_hw)


class Siddon(torch.nn.Module):
    """Exact voxel-traversal renderer (``diffdrr.renderers.Siddon``)."""

    def __init__(self, voxel_shift=conv.SIDDON_VOXEL_SHIFT_DEFAULT, mode="nearest",
                 stop_gradients_through_grid_sample=False, filter_intersections_outside_volume=True,
                 reducefn="sum", eps=conv.RENDER_EPS):
        super().__init__()
        if mode != "nearest" or not filter_intersections_outside_volume or reducefn != "sum":
            raise NotImplementedError(
                "xvr_b200.Siddon implements mode='nearest', filter_intersections_outside_volume=True, "
                "reducefn='sum'")
        self.voxel_shift = voxel_shift
        self.eps = eps
        self.detector_hw = None
        self._labels = _LabelCache()

    def dims(self, volume):
        return torch.tensor(volume.shape).to(volume) + 1

    def forward(self, volume, source, target, img, mask=None):
        from ._siddon import siddon_rays  # noqa: PLC0415  (kept separate: its library half is built -fmad=false)

        labels, C = (None, 1) if mask is None else self._labels.get(mask)
        return siddon_rays(volume, source, target, img, labels, C, float(self.voxel_shift), float(self.eps),
                           self.detector_hw)
