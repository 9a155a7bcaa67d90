"""Ray-march renderers: drop-in for ``diffdrr.renderers.{Trilinear, Siddon}`` on hand-written sm_100a kernels.

Call signature pinned by /root/reference/src/xvr/model/trainer.py:288:
``renderer(volume (D0,D1,D2), source (B,1,3), target (B,N,3), raylen (B,1,N), mask=labelmap|None) -> (B,C,N)``
with ``source``/``target`` in voxel-index coordinates.  The modules are autograd-transparent: gradients flow
to ``source``, ``target``, ``raylen`` (pose path) and, on request, to ``volume``.
"""

import ctypes
import os
import weakref

import torch

from . import _conventions as conv
from . import _lib
from ._lib import call, cuda_f32, ptr, stream

__all__ = ["Trilinear", "Siddon"]


def _tile_shape(kind="trilinear"):
    """(lane_w_log2, cta_w_log2), unless overridden for tuning.  Trilinear: each warp takes a 32-row x 1-column strip of
    detector pixels and a CTA eight adjacent strips (detector rows run along the volume's contiguous axis in the usual
    AP/PA set-up: the TLD4 footprints of neighbouring lanes overlap).  Siddon: compact 4-row x 8-column warps, eight
    stacked in a CTA -- its scattered one-voxel gathers care less about the axis than about rays of similar length
    sharing a warp (measured on config 5: 33.9 vs 35.7 ms at B = 64, scripts/sweep_tiles.py)."""
    env = os.environ.get("XVR_B200_TILE")
    if env:
        lw, cw = (int(v) for v in env.split(","))
        return lw, cw
    return (3, 3) if kind == "siddon" else (0, 3)


class _VolumeTexture:
    """Block-linear (layered cudaArray) copy of the volume behind a texture object, re-uploaded whenever the
    source tensor is a different object or has been modified in place (torch's version counter)."""

    def __init__(self, texture=True):
        self.texture = texture  # False: occupancy only (xvr_occupancy_create) -- what the Siddon entries take
        self.handle = None
        self.shape = None
        self.device = None
        self.src = None  # weakref to the tensor last uploaded
        self.version = None

    def get(self, volume):
        mode = os.environ.get("XVR_B200_GATHER", "tex")
        if mode == "ldg" and self.texture:
            return None
        shape = tuple(volume.shape)
        if self.handle is None or shape != self.shape or volume.device != self.device:
            self.free()
            h = ctypes.c_void_p()
            with torch.cuda.device(volume.device):
                call("xvr_volume_create" if self.texture else "xvr_occupancy_create", *shape, ctypes.byref(h))
            self.handle, self.shape, self.device, self.src = h, shape, volume.device, None
        if self.src is None or self.src() is not volume or self.version != volume._version:
            call("xvr_volume_upload", self.handle, ptr(volume), stream())
            self.src, self.version = weakref.ref(volume), volume._version
        return self.handle

    def invalidate(self):
        """Forget what was uploaded last: the next ``get`` re-uploads.  Needed when the source tensor is rewritten
        through its raw pointer (no version bump), e.g. the static density buffer of a graph-captured step."""
        self.src = None

    def __deepcopy__(self, memo):
        return _VolumeTexture(self.texture)  # device resources are per-instance: the copy re-creates its own lazily

    def free(self):
        if self.handle is not None:
            try:
                _lib.lib().xvr_volume_destroy(self.handle)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self.handle = None

    def __del__(self):
        self.free()


class _LabelCache:
    """uint8 copies of label volumes and their channel counts, one entry per source tensor and version.  Several
    entries stay alive at once: a training run alternates between subjects, and CUDA graphs captured for one
    subject keep raw pointers into its entry.

    (Measured and dropped: the label map as a second layered texture -- the extra point fetch per sample costs the
    saturated texture pipe more than the scattered 1-byte load costs the LSU path: 19.2 -> 20.7 ms at C2.)"""

    MAX_ENTRIES = 32

    def __init__(self):
        self.entries = {}  # key -> (labels uint8, channels); insertion order = age

    def __deepcopy__(self, memo):
        return _LabelCache()

    def get(self, mask):
        key = (mask.data_ptr(), mask._version, tuple(mask.shape), mask.dtype, mask.device)
        hit = self.entries.get(key)
        # The key alone is not an identity: the reference's Trainer.load builds a fresh float mask every iteration
        # (`mask.to(device, dtype)`), the caching allocator hands the freed block back, and the new tensor has the
        # old pointer, shape and version 0.  An entry is a hit only while its SOURCE TENSOR OBJECT is alive and is
        # the one asked about (the uint8 copy is kept alive by the entry, the source only weakly).
        if hit is not None and hit[2]() is not mask:
            self.entries.pop(key)
            hit = None
        if hit is None:
            hi = int(mask.max().item())  # same host sync as the reference's `int(mask.max()) + 1`
            lo = int(mask.min().item())
            if lo < 0 or hi > 254:
                raise _lib.XvrB200Error(f"label volume must hold integers in [0, 254]; got [{lo}, {hi}]")
            for k in [k for k, v in self.entries.items() if v[2]() is None]:
                self.entries.pop(k)  # sources that died
            while len(self.entries) >= self.MAX_ENTRIES:
                self.entries.pop(next(iter(self.entries)))
            hit = (_labels_with_brick_table(mask), hi + 1, weakref.ref(mask))
            self.entries[key] = hit
        return hit[0], hit[1]


LABEL_BRICK = 8  # = OCC_BRICK of csrc/common.cuh
OPT_LABEL_BRICKS = 0x80


def _labels_with_brick_table(mask):
    """The uint8 label volume followed -- at the next multiple of 256 bytes -- by the brick table of
    XVR_OPT_LABEL_BRICKS (include/xvr_b200.h): per 8^3 brick the label all voxels of the brick grown by one voxel carry
    (outside the volume = 0), 255 if they differ.  Set-up work on the label map, once per mask object (like the uint8
    conversion itself); the trilinear forward answers the nearest-label lookup of uniform bricks from the table.
    Returns the (D0,D1,D2) uint8 label volume as a view of the first D0*D1*D2 bytes of that buffer (what every kernel
    reads as `labels`; `_has_brick_table` tells the two apart)."""
    lab = mask.to(torch.uint8).contiguous()
    n = lab.numel()
    nb = [(d + LABEL_BRICK - 1) // LABEL_BRICK for d in lab.shape]
    # min / max over every brick grown by one voxel: pad by 1 in front and up to the brick grid + 1 behind with 0 (= the
    # label of everything outside), pool with window 10 / stride 8.  fp16 holds 0..255 exactly.
    pad = []
    for d, b in zip(reversed(lab.shape), reversed(nb)):
        pad += [1, b * LABEL_BRICK + 1 - d]
    x = torch.nn.functional.pad(lab.to(torch.float16)[None, None], pad)
    hi = torch.nn.functional.max_pool3d(x, LABEL_BRICK + 2, LABEL_BRICK)
    lo = -torch.nn.functional.max_pool3d(-x, LABEL_BRICK + 2, LABEL_BRICK)
    table = torch.where(hi == lo, hi, torch.full_like(hi, 255.0)).to(torch.uint8).reshape(-1)
    assert table.numel() == nb[0] * nb[1] * nb[2]
    off = (n + 255) // 256 * 256
    buf = torch.zeros(off + table.numel(), dtype=torch.uint8, device=lab.device)
    buf[:n] = lab.reshape(-1)
    buf[off:] = table
    return buf[:n].view(lab.shape)


def _has_brick_table(labels):
    return labels.untyped_storage().nbytes() > labels.numel()


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


class _RenderRays(torch.autograd.Function):
    """(volume, source, target, raylen) -> (B,C,N) through the C-ABI; ``kind`` selects the renderer and ``args``
    holds its scalar arguments in C-ABI order (trilinear: n_points, step_mode, eps; siddon: voxel_shift, eps).

    When pose gradients are needed and there are no label channels, the forward kernel also emits the per-ray
    Jacobian d out / d(source, target, raylen) -- it does not depend on the upstream gradient -- so the backward
    pass is a 28-byte-per-ray epilogue instead of a second march through the volume.
    """

    @staticmethod
    def forward(ctx, volume, source, target, raylen, labels, C, kind, args, det_hw, voltex):
        source, target, raylen = cuda_f32(source, "source"), cuda_f32(target, "target"), cuda_f32(raylen, "raylen")
        B, N = _check_rays(volume, source, target, raylen)
        lw, cw = _tile_shape(kind)
        det_h, det_w = det_hw if det_hw is not None and det_hw[0] * det_hw[1] == N else (0, 0)
        need_pose_grad = any(ctx.needs_input_grad[1:4])
        need_vol_grad = ctx.needs_input_grad[0]
        out = torch.empty(B, C, N, device=volume.device, dtype=torch.float32)
        ctx.empty = B == 0 or N == 0
        if ctx.empty:  # empty in -> empty out, as grid_sample does; nothing to launch
            return out
        # with label channels the Jacobian is the one of the channel SUM: enough whenever the caller collapses the
        # channels (trainer.py:294 img.sum(dim=1)); backward() falls back to the recompute kernel otherwise
        jac = (torch.empty(B, 7, N, device=volume.device, dtype=torch.float32)
               if need_pose_grad and not need_vol_grad else None)  # d/dvolume re-marches anyway: no Jacobian to keep
        vol_args = (ptr(volume),) if voltex is False else (ptr(volume), voltex)
        ctx.common = (*vol_args, *volume.shape, ptr(labels), C, ptr(source), ptr(target), ptr(raylen), B, N, *args,
                      det_h, det_w, lw, cw)
        # the label buffers of _LabelCache carry their brick table (XVR_OPT_LABEL_BRICKS; trilinear forward only)
        bricks = (OPT_LABEL_BRICKS if labels is not None and kind == "trilinear"
                  and _has_brick_table(labels) and os.environ.get("XVR_B200_LABEL_BRICKS", "1") != "0" else 0)
        call(f"xvr_{kind}_rays_fwd", *ctx.common, ptr(out), ptr(jac), _lib.opts_word() | bricks, stream())
        ctx.kind, ctx.C, ctx.shape = kind, C, tuple(volume.shape)
        if need_pose_grad or need_vol_grad:
            # the recompute path (per-channel upstream gradients, d/dvolume) needs the inputs; saving them keeps the
            # raw pointers of ctx.common alive
            keep = labels is not None or need_vol_grad
            ctx.save_for_backward(jac, *((volume, source, target, raylen, labels) if keep else ()))
        return out

    @staticmethod
    def backward(ctx, gout):
        need_vol_grad = ctx.needs_input_grad[0]
        gout_shared = (gout.shape[1] == 1 or gout.stride(1) == 0) and not need_vol_grad
        gout = cuda_f32(gout[:, :1] if gout_shared else gout, "grad_output")
        B, _, N = gout.shape
        dev = gout.device
        gsource = torch.empty(B, 1, 3, device=dev, dtype=torch.float32)
        gtarget = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        graylen = torch.empty(B, 1, N, device=dev, dtype=torch.float32)
        # d/dvolume from the ray entry point: the reference's own formulation (one RED.ADD per corner / segment into a
        # zeroed volume, non-deterministic summation order) -- rays given as tensors carry no detector geometry to
        # derive an atomics-free ownership from; DRR.forward's fused path has one (csrc/volgrad.cu, siddon_volgrad.cu)
        gvol = torch.zeros(ctx.shape, device=dev, dtype=torch.float32) if need_vol_grad else None
        if ctx.empty:
            return gvol, gsource.zero_(), gtarget, graylen, None, None, None, None, None, None
        work = torch.empty(B, 3, N, device=dev, dtype=torch.float32)
        jac = ctx.saved_tensors[0]
        # One upstream gradient per ray: single channel, or every channel sees the same gradient -- autograd hands
        # the backward of sum(dim=1) over as an expanded view (stride 0 along the channel axis), a host-side check.
        if gout_shared:
            call("xvr_rays_jac_bwd", ptr(jac), ptr(gout), B, N, ptr(gsource), ptr(gtarget), ptr(graylen),
                 ptr(work), stream())
        else:
            call(f"xvr_{ctx.kind}_rays_bwd", *ctx.common, ptr(gout), ptr(gsource), ptr(gtarget), ptr(graylen),
                 ptr(work), ptr(gvol), *((_lib.opts_word(),) if ctx.kind == "siddon" else ()), stream())
        return gvol, gsource, gtarget, graylen, None, None, None, None, None, None


# Diagnostics of the staged variant: set _staged_stats["tensor"] to a zeroed int64 CUDA tensor of 3 elements to collect
# {samples served from shared memory, from global memory, barrier time-outs}.
_staged_stats = {}


class _RenderDRR(torch.autograd.Function):
    """Fused DRR: rays are generated inside the kernel from cam2vox (B,3,4) and the detector basis, so no
    (B,N,3) tensor exists; the backward reduces the saved per-ray Jacobian straight to dL/dcam2vox (B,3,4) and,
    if the volume requires a gradient, gathers dL/dvolume voxel by voxel (no atomics, deterministic)."""

    @staticmethod
    def forward(ctx, volume, cam2vox, cam2world, det9, det_hw, kind, args, voltex, labels=None, C=1):
        # voltex: the volume's texture handle (trilinear) / occupancy handle (siddon), or None
        # labels / C: uint8 label volume of _LabelCache and its channel count (trilinear only) -> out (B,C,H*W)
        cam2vox, cam2world = cuda_f32(cam2vox, "cam2vox"), cuda_f32(cam2world, "cam2world")
        B = cam2vox.shape[0]
        H, W = det_hw
        lw, cw = _tile_shape(kind)
        out = torch.empty(B, C, H * W, device=volume.device, dtype=torch.float32)
        jac = torch.empty(B, 7, H * W, device=volume.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        det = (ctypes.c_float * 9)(*det9)
        ctx.det = (det, B, H, W, args, tuple(volume.shape))
        ctx.kind, ctx.labels, ctx.C, ctx.det9, ctx.volume = kind, labels, C, det9, None
        if B == 0:  # an empty pose batch renders to an empty image batch; nothing to launch
            ctx.save_for_backward(jac)
            return out
        staged = os.environ.get("XVR_B200_STAGED", "0")
        if labels is not None:
            if kind != "trilinear" or ctx.needs_input_grad[0]:
                raise _lib.XvrB200Error("fused label channels: trilinear renderer, no volume gradient")
            bricks = (OPT_LABEL_BRICKS if _has_brick_table(labels)
                      and os.environ.get("XVR_B200_LABEL_BRICKS", "1") != "0" else 0)
            call("xvr_trilinear_drr_fwd_labels", ptr(volume), voltex, *volume.shape, ptr(labels), C, ptr(cam2vox),
                 ptr(cam2world), det, B, H, W, *args, lw, cw, ptr(out), ptr(jac), _lib.opts_word() | bricks, stream())
            ctx.volume = volume  # the per-channel-gradient fallback of backward() renders again
            ctx.save_for_backward(jac, cam2vox, cam2world)
            return out
        if kind == "siddon":
            call("xvr_siddon_drr_fwd", ptr(volume), voltex, *volume.shape, ptr(cam2vox), ptr(cam2world), det, B, H, W, *args,
                 lw, cw, ptr(out), ptr(jac), _lib.opts_word(), stream())
        elif staged == "1" and volume.shape[2] % 4 == 0 and volume.data_ptr() % 16 == 0:
            # bricks staged in shared memory by the TMA unit (csrc/trilinear_staged.cu)
            call("xvr_trilinear_drr_fwd_staged", ptr(volume), *volume.shape, ptr(cam2vox), ptr(cam2world), det, B, H, W,
                 *args, ptr(out), ptr(jac), ptr(_staged_stats.get("tensor")), stream())
        else:
            call("xvr_trilinear_drr_fwd", ptr(volume), voltex, *volume.shape, ptr(cam2vox), ptr(cam2world), det, B, H,
                 W, *args, lw, cw, ptr(out), ptr(jac), _lib.opts_word(), stream())
        ctx.det = (det, B, H, W, args, tuple(volume.shape))
        ctx.kind = kind
        ctx.save_for_backward(jac, *((cam2vox, cam2world) if ctx.needs_input_grad[0] else ()))
        return out

    @staticmethod
    def backward(ctx, gout):
        jac, *mats = ctx.saved_tensors
        det, B, H, W, args, shape = ctx.det
        # one upstream gradient for every channel?  (autograd hands the backward of sum(dim=1) over as a stride-0
        # expansion: checked on the tensor as it arrives, before it is made contiguous)
        shared = gout.shape[1] == 1 or gout.stride(1) == 0
        if ctx.labels is not None and shared:
            gout = gout[:, :1]
        gout = cuda_f32(gout, "grad_output")
        gG = gvol = None
        if B == 0:
            return (torch.zeros(shape, device=gout.device) if ctx.needs_input_grad[0] else None,
                    torch.zeros(0, 3, 4, device=gout.device) if ctx.needs_input_grad[1] else None,
                    None, None, None, None, None, None, None, None)
        if ctx.labels is not None and ctx.needs_input_grad[1]:
            # label channels: the saved Jacobian is the one of the channel SUM -- exact whenever every channel sees the
            # same upstream gradient (autograd hands the backward of sum(dim=1) over as a stride-0 expansion).
            if not shared:
                return (None, _labelled_drr_backward_by_rays(ctx, mats[0], mats[1], gout, H, W, args),
                        None, None, None, None, None, None, None, None)
        if ctx.needs_input_grad[1]:
            gG = torch.empty(B, 3, 4, device=gout.device, dtype=torch.float32)
            slices = _lib.lib().xvr_drr_jac_bwd_slices(B, H * W)
            work = torch.empty(12 * B * slices, device=gout.device, dtype=torch.float32) if slices > 1 else None
            call("xvr_drr_jac_bwd", ptr(jac), ptr(gout), det, B, H, W, ptr(gG), ptr(work), stream())
        if ctx.needs_input_grad[0]:
            cam2vox, cam2world = mats
            full = torch.zeros(B, 4, 4, device=gout.device, dtype=torch.float64)
            full[:, :3] = cam2vox
            full[:, 3, 3] = 1.0
            vox2cam = torch.linalg.inv(full)[:, :3].to(torch.float32).contiguous()
            gvol = torch.empty(shape, device=gout.device, dtype=torch.float32)
            if ctx.kind == "siddon":
                call("xvr_siddon_drr_bwd_volume", ptr(cam2vox), ptr(vox2cam), ptr(cam2world), det, B, H, W, *args,
                     ptr(gout), *shape, ptr(gvol), 0, _lib.opts_word(), stream())
            else:
                work = torch.empty(B * H * W * 12, device=gout.device, dtype=torch.float32)
                call("xvr_trilinear_drr_bwd_volume", ptr(cam2vox), ptr(vox2cam), ptr(cam2world), det, B, H, W, *args,
                     ptr(gout), *shape, ptr(work), ptr(gvol), 0, _lib.opts_word(), stream())
        return gvol, gG, None, None, None, None, None, None, None, None


def _labelled_drr_backward_by_rays(ctx, cam2vox, cam2world, gout, H, W, args):
    """dL/dcam2vox of the fused label-channel render for an upstream gradient that differs between channels (rare: xvr
    collapses the channels before it differentiates): the rays are materialised from the camera matrices and the
    ray entry point's recompute backward does the work, autograd chaining it back to the matrices."""
    o, u, v = (torch.tensor(ctx.det9[3 * i:3 * i + 3], device=gout.device, dtype=torch.float32) for i in range(3))
    ii = torch.arange(H, device=gout.device, dtype=torch.float32).repeat_interleave(W)
    jj = torch.arange(W, device=gout.device, dtype=torch.float32).repeat(H)
    pts = o + ii[:, None] * u + jj[:, None] * v  # (N,3) detector points in the camera frame, row-major
    with torch.enable_grad():
        G = cam2vox.detach().requires_grad_()
        source = G[:, :, 3].unsqueeze(1)
        target = torch.einsum("bij,nj->bni", G[:, :, :3], pts) + source
        world = torch.einsum("bij,nj->bni", cam2world[:, :, :3], pts)  # target - source in world mm
        raylen = world.norm(dim=-1).unsqueeze(1)
        out = _RenderRays.apply(ctx.volume, source.contiguous(), target.contiguous(), raylen.contiguous(), ctx.labels,
                                ctx.C, "trilinear", args, (H, W), None)
        (gG,) = torch.autograd.grad(out, G, gout)
    return gG


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
        self._texture = _VolumeTexture()

    def dims(self, volume):
        return torch.tensor(volume.shape).to(volume) - 1

    def forward(self, volume, source, target, img, n_points=conv.TRILINEAR_N_POINTS, align_corners=True,
                mask=None):
        if not align_corners:
            raise NotImplementedError("xvr_b200.Trilinear implements align_corners=True (the DiffDRR default)")
        labels, C = (None, 1) if mask is None else self._labels.get(mask)
        volume = cuda_f32(volume, "volume")
        return _RenderRays.apply(volume, source, target, img, labels, C, "trilinear",
                                 (int(n_points), conv.STEP_MODES[self.step], float(self.eps)), self.detector_hw,
                                 self._texture.get(volume))


    def render_drr(self, volume, cam2vox, cam2world, detector, n_points=conv.TRILINEAR_N_POINTS, mask=None):
        """Fused path used by ``DRR.forward``: ``cam2vox``/``cam2world`` are (B,3,4) camera->voxel / camera->world
        matrices, ``detector`` supplies the pixel grid.  Returns (B,1,H*W), or (B,C,H*W) with a label map ``mask``."""
        volume = cuda_f32(volume, "volume")
        origin, row_step, col_step = detector.pixel_basis()
        labels, C = (None, 1) if mask is None else self._labels.get(mask)
        return _RenderDRR.apply(volume, cam2vox, cam2world, (*origin, *row_step, *col_step),
                                (detector.height, detector.width), "trilinear",
                                (int(n_points), conv.STEP_MODES[self.step], float(self.eps)),
                                self._texture.get(volume), labels, C)


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
        self._texture = _VolumeTexture(texture=False)  # occupancy only: the traversal gathers from the linear volume

    def dims(self, volume):
        return torch.tensor(volume.shape).to(volume) + 1

    def forward(self, volume, source, target, img, mask=None):
        labels, C = (None, 1) if mask is None else self._labels.get(mask)
        return _RenderRays.apply(cuda_f32(volume, "volume"), source, target, img, labels, C, "siddon",
                                 (float(self.voxel_shift), float(self.eps)), self.detector_hw, False)

    def render_drr(self, volume, cam2vox, cam2world, detector):
        """Fused path used by ``DRR.forward``: rays are generated inside the kernel (no (B,N,3) tensors)."""
        volume = cuda_f32(volume, "volume")
        origin, row_step, col_step = detector.pixel_basis()
        return _RenderDRR.apply(volume, cam2vox, cam2world, (*origin, *row_step, *col_step),
                                (detector.height, detector.width), "siddon",
                                (float(self.voxel_shift), float(self.eps)), self._texture.get(volume), None, 1)
