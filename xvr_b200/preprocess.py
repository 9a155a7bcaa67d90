"""X-ray intensity transforms that sit between the renderer and the similarity metric in both hot loops
(/root/reference/src/xvr/utils/preprocess.py:5-66, used at model/trainer.py:207,216 and
registrar/base.py:213-218,250): batch-global min/max standardisation, optional soft-histogram equalisation,
resize to the working resolution, and (x - mean) / std."""

import torch
import torch.nn.functional as F

__all__ = ["XrayTransforms", "Standardize", "Equalize"]


class Standardize(torch.nn.Module):
    """(x - min) / (max - min + eps) with min/max taken over the WHOLE batch tensor."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        lo, hi = x.min(), x.max()  # differentiable (aminmax is not)
        return (x - lo) / (hi - lo + self.eps)


class Equalize(torch.nn.Module):
    """Differentiable histogram equalisation with a Gaussian soft histogram (256 bins, tau = 0.01)."""

    def __init__(self, n_bins=256, tau=0.01, eps=1e-10):
        super().__init__()
        self.n_bins, self.tau, self.eps = n_bins, tau, eps

    def forward(self, x):
        B, _, H, W = x.shape
        centres = torch.linspace(0, 1, self.n_bins, device=x.device)
        w = torch.exp(-(x.reshape(B, -1, 1) - centres).square() / (2 * self.tau**2))  # (B, HW, bins)
        hist = w.sum(1)
        hist = hist / (hist.sum(1, keepdim=True) + self.eps)
        cdf = hist.cumsum(1)
        cdf = (cdf - cdf[:, :1]) / (1 - cdf[:, :1] + self.eps)
        w = w / (w.sum(-1, keepdim=True) + self.eps)
        return (w * cdf[:, None]).sum(-1).view(B, 1, H, W)


class XrayTransforms(torch.nn.Module):
    """Standardize -> [Equalize] -> Resize((height, width)) -> Normalize(mean, std)."""

    def __init__(self, height, width=None, mean=0.15, std=0.1, equalize=False):
        super().__init__()
        self.size = (int(height), int(height if width is None else width))
        self.mean, self.std = mean, std
        self.standardize = Standardize()
        self.equalize = Equalize() if equalize else None

    def forward(self, x):
        x = self.standardize(x)
        if self.equalize is not None:
            x = self.equalize(x)
        if tuple(x.shape[-2:]) != self.size:
            # torchvision.transforms.Resize on tensors: bilinear, antialiased
            x = F.interpolate(x, size=self.size, mode="bilinear", align_corners=False, antialias=True)
        return (x - self.mean) / self.std
