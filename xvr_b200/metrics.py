"""Image similarity and pose distances: drop-in for the slice of ``diffdrr.metrics`` xvr uses.

``MultiscaleNormalizedCrossCorrelation2d([None, p], [0.5, 0.5])`` (/root/reference/src/xvr/model/loss.py:16,27;
registrar/base.py:119-121), ``GradientNormalizedCrossCorrelation2d(patch_size, sigma)`` (registrar/base.py:122)
and ``DoubleGeodesicSE3(sdd)`` (model/loss.py:18,29,46; metrics/evaluator.py:15,34).  The NCC family runs in the
kernels of csrc/ncc.cu, forward and backward; the geodesic is O(B) arithmetic on 3x3 matrices and stays in
PyTorch.
"""

import torch

from . import _conventions as conv
from . import _lib
from ._lib import call, cuda_f32, ptr, stream
from .pose import so3_log_map

__all__ = [
    "RegistrationSimilarity",
    "NormalizedCrossCorrelation2d",
    "MultiscaleNormalizedCrossCorrelation2d",
    "GradientNormalizedCrossCorrelation2d",
    "Sobel",
    "DoubleGeodesicSE3",
]

_NCC_T = 16


def _coef_shape(B, C, H, W, p):
    return (B * C, 6) if p is None else (B * C, 4, H - p + 1, W - p + 1)


class _MultiNCC(torch.autograd.Function):
    """score (B,) = sum_k w_k * NCC_{p_k}(x1, x2); p_k = None is the whole-image NCC."""

    @staticmethod
    def forward(ctx, x1, x2, patches, weights, eps):
        x1, x2 = cuda_f32(x1, "x1"), cuda_f32(x2, "x2")
        if x1.shape != x2.shape or x1.dim() != 4:
            raise ValueError(f"expected two (B,C,H,W) images of equal shape; got {tuple(x1.shape)}, {tuple(x2.shape)}")
        B, C, H, W = x1.shape
        score = torch.empty(B, device=x1.device, dtype=torch.float32)
        if B == 0:  # empty batch -> empty score vector (the reference's mean over dims 1.. of an empty batch)
            ctx.cfg = None
            return score
        tiles = max(((H + _NCC_T - 1) // _NCC_T) * ((W + _NCC_T - 1) // _NCC_T), 6)
        work = torch.empty(B * C * tiles, device=x1.device, dtype=torch.float32)
        g1, g2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        saved = []
        for k, (p, w) in enumerate(zip(patches, weights)):
            c2 = torch.empty(_coef_shape(B, C, H, W, p), device=x1.device) if g2 else None
            c1 = torch.empty(_coef_shape(B, C, H, W, p), device=x1.device) if g1 else None
            call("xvr_ncc_fwd", ptr(x1), ptr(x2), B, C, H, W, 0 if p is None else int(p), float(eps), float(w),
                 int(k > 0), ptr(score), ptr(work), ptr(c2), ptr(c1), stream())
            saved += [c2, c1]
        ctx.cfg = (patches, weights)
        ctx.save_for_backward(x1, x2, *saved)
        return score

    @staticmethod
    def backward(ctx, gscore):
        if ctx.cfg is None:
            return None, None, None, None, None
        x1, x2, *saved = ctx.saved_tensors
        patches, weights = ctx.cfg
        B, C, H, W = x1.shape
        gscore = cuda_f32(gscore, "grad_output")
        grads = [None, None]
        for which, slot in ((1, 0), (2, 1)):
            if not ctx.needs_input_grad[slot]:
                continue
            g = torch.empty_like(x1)
            for k, (p, w) in enumerate(zip(patches, weights)):
                coef = saved[2 * k + (0 if which == 2 else 1)]
                call("xvr_ncc_bwd", ptr(x1), ptr(x2), ptr(coef), which, ptr(gscore), B, C, H, W,
                     0 if p is None else int(p), float(w), int(k > 0), ptr(g), stream())
            grads[slot] = g
        return grads[0], grads[1], None, None, None


class _Sobel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = cuda_f32(x, "image")
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"Sobel expects (B,1,H,W); got {tuple(x.shape)}")
        B, _, H, W = x.shape
        out = torch.empty(B, 2, H, W, device=x.device, dtype=torch.float32)
        if B > 0:
            call("xvr_sobel_fwd", ptr(x), B, H, W, ptr(out), stream())
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = cuda_f32(gout, "grad_output")
        B, _, H, W = gout.shape
        gx = torch.empty(B, 1, H, W, device=gout.device, dtype=torch.float32)
        if B > 0:
            call("xvr_sobel_bwd", ptr(gout), B, H, W, ptr(gx), stream())
        return gx


class _RegSim(torch.autograd.Function):
    """moving (B,1,H,W) raw DRRs -> scalar similarity to a fixed, transformed X-ray, value AND gradient in one C-ABI
    call (nine launches, csrc/ncc.cu xvr_regsim); backward scales the saved gradient by the upstream scalar."""

    @staticmethod
    def forward(ctx, moving, fixed, fixed_sobel, cfg):
        moving = cuda_f32(moving, "moving image")
        if moving.shape != fixed.shape or moving.dim() != 4 or moving.shape[1] != 1:
            raise ValueError(f"expected (B,1,H,W) images of equal shape; got {tuple(moving.shape)}, {tuple(fixed.shape)}")
        B, _, H, W = moving.shape
        score = torch.zeros((), device=moving.device, dtype=torch.float32)
        grad = torch.zeros_like(moving)
        if B > 0:
            std_eps, mean, inv_std, p, q, w_global, w_patch, w_grad, eps = cfg
            n = _lib.lib().xvr_regsim_workspace_floats(B, H, W, p, q)
            if n < 0:
                raise _lib.XvrB200Error(f"fused registration similarity does not support B={B}, {H}x{W}, patches {p}, {q}")
            work = torch.empty(n, device=moving.device, dtype=torch.float32)
            call("xvr_regsim", ptr(fixed), ptr(fixed_sobel), ptr(moving), B, H, W, std_eps, mean, inv_std, p, q,
                 w_global, w_patch, w_grad, eps, ptr(work), n, ptr(score), ptr(grad), stream())
        ctx.save_for_backward(grad)
        return score

    @staticmethod
    def backward(ctx, gscore):
        (grad,) = ctx.saved_tensors
        return grad * gscore, None, None, None


class RegistrationSimilarity(torch.nn.Module):
    """The similarity one registration iteration maximises, from the RAW moving DRR to the scalar:

        ``(beta * mNCC([None, p], [0.5, 0.5]) + (1 - beta) * GradNCC(q, sigma=0))(fixed, XrayTransforms(moving)).sum()``

    (/root/reference/src/xvr/registrar/base.py:119-122, 245-252) as one fused value-and-gradient evaluation.
    ``fixed`` is the already transformed target; its Sobel image is computed once here.  Covers the registration
    defaults only: no ``Equalize``, no resize, no Gaussian blur in front of the Sobel operator."""

    def __init__(self, fixed, mncc_patch_size=9, gncc_patch_size=11, beta=0.5, mean=0.15, std=0.1, std_eps=1e-6,
                 eps=conv.NCC_EPS):
        super().__init__()
        fixed = cuda_f32(fixed.detach(), "fixed image")
        self.register_buffer("fixed", fixed, persistent=False)
        self.register_buffer("fixed_sobel", _Sobel.apply(fixed), persistent=False)
        # x / std on a CUDA tensor multiplies by the fp32 reciprocal of the Python scalar (ATen div_true_kernel_cuda)
        inv_std = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(std, dtype=torch.float32))
        self.cfg = (float(std_eps), float(mean), inv_std, int(mncc_patch_size), int(gncc_patch_size), 0.5 * beta,
                    0.5 * beta, 1.0 - beta, float(eps))

    def forward(self, moving):
        return _RegSim.apply(moving, self.fixed, self.fixed_sobel, self.cfg)


class NormalizedCrossCorrelation2d(torch.nn.Module):
    """Mean over channels (and windows) of the z-scored correlation; ``patch_size=None`` is the global NCC."""

    def __init__(self, patch_size=None, eps=conv.NCC_EPS):
        super().__init__()
        self.patch_size = patch_size
        self.eps = eps

    def forward(self, x1, x2):
        return _MultiNCC.apply(x1, x2, (self.patch_size,), (1.0,), self.eps)


class MultiscaleNormalizedCrossCorrelation2d(torch.nn.Module):
    """Weighted sum of NCCs at several patch sizes, e.g. ``([None, 9], [0.5, 0.5])``."""

    def __init__(self, patch_sizes=(None,), patch_weights=(1.0,), eps=conv.NCC_EPS):
        super().__init__()
        if len(patch_sizes) != len(patch_weights):
            raise ValueError("patch_sizes and patch_weights must have the same length")
        self.patch_sizes = tuple(patch_sizes)
        self.patch_weights = tuple(float(w) for w in patch_weights)
        self.eps = eps

    def forward(self, x1, x2):
        return _MultiNCC.apply(x1, x2, self.patch_sizes, self.patch_weights, self.eps)


class Sobel(torch.nn.Module):
    """3x3 Sobel gradients (B,1,H,W) -> (B,2,H,W); a Gaussian pre-blur is applied only when sigma > 0."""

    def __init__(self, sigma=0.0):
        super().__init__()
        self.sigma = float(sigma)
        # kept as buffers so that .cuda()/.to() on the owning metric behaves like the reference's Conv2d filter
        self.register_buffer("filter", torch.tensor(
            [[[[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]]],
             [[[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]]]]))

    def forward(self, img):
        if self.sigma > 0:
            k = torch.arange(5, dtype=img.dtype, device=img.device) - 2
            g = torch.exp(-0.5 * (k / self.sigma) ** 2)
            g = g / g.sum()
            pad = torch.nn.functional.pad(img, (2, 2, 2, 2), mode="reflect")
            img = torch.nn.functional.conv2d(torch.nn.functional.conv2d(pad, g.view(1, 1, 1, 5)), g.view(1, 1, 5, 1))
        return _Sobel.apply(img)


class GradientNormalizedCrossCorrelation2d(torch.nn.Module):
    """NCC of the Sobel gradient images."""

    def __init__(self, patch_size=None, sigma=1.0, eps=conv.NCC_EPS):
        super().__init__()
        self.patch_size = patch_size
        self.eps = eps
        self.sobel = Sobel(sigma)

    def forward(self, x1, x2):
        return _MultiNCC.apply(self.sobel(x1), self.sobel(x2), (self.patch_size,), (1.0,), self.eps)


class DoubleGeodesicSE3(torch.nn.Module):
    """(angular, translational, double) geodesic distances between two poses; angular = sdd/2 * |log(R1^T R2)|."""

    def __init__(self, sdd, eps=conv.GEODESIC_EPS):
        super().__init__()
        self.sdr = sdd / 2.0
        self.eps = eps

    def forward(self, pose_1, pose_2):
        r = pose_1.rotation.transpose(-1, -2) @ pose_2.rotation
        angular = self.sdr * so3_log_map(r).norm(dim=-1)
        translational = (pose_1.translation - pose_2.translation).norm(dim=-1)
        return angular, translational, (angular.square() + translational.square() + self.eps).sqrt()
