"""``diffdrr.visualization.{plot_drr, plot_mask}`` -- import compatibility for xvr (model/trainer.py:8,
registrar/base.py:12, visualization/*.py); plotting is outside the hot path (OUT OF SCOPE).

Minimal matplotlib versions are provided so that xvr's logging / ``--verbose 3`` code paths work when matplotlib is
installed; without it they raise an ImportError that says so (the build container has no matplotlib).
"""

import torch

__all__ = ["plot_drr", "plot_mask"]


def _pyplot():
    try:
        import matplotlib.pyplot as plt  # noqa: PLC0415
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("xvr_b200.visualization needs matplotlib, which is not installed") from e
    return plt


def _axes(n, axs):
    plt = _pyplot()
    if axs is None:
        _, axs = plt.subplots(ncols=n, figsize=(3 * n, 3))
    return [axs] if n == 1 and not isinstance(axs, (list, tuple)) and not hasattr(axs, "__len__") else list(axs)


def plot_drr(img, title=None, ticks=True, axs=None, cmap="gray", **imshow_kwargs):
    """Show a batch of DRRs (B,C,H,W); multi-channel images are summed over channels, as DiffDRR does."""
    img = img.detach().sum(dim=1, keepdim=True).cpu() if img.shape[1] > 1 else img.detach().cpu()
    axs = _axes(len(img), axs)
    titles = title if isinstance(title, (list, tuple)) else [title] * len(img)
    for x, ax, t in zip(img, axs, titles):
        ax.imshow(x.squeeze(), cmap=cmap, **imshow_kwargs)
        _, height, width = x.shape
        ax.xaxis.tick_top()
        ax.set(title=t, xticks=[0, width - 1], xticklabels=[1, width], yticks=[0, height - 1], yticklabels=[1, height])
        if not ticks:
            ax.set_xticks([])
            ax.set_yticks([])
    return axs


def plot_mask(img, axs=None, colors=None, alpha=0.5, return_masks=False):
    """Overlay the channels of a multi-channel mask (B,C,H,W), one colour per channel."""
    plt = _pyplot()
    img = img.detach().cpu()
    B, C, H, W = img.shape
    axs = _axes(B, axs)
    cmap = plt.get_cmap("rainbow", max(C, 1)) if colors is None else None
    masks = torch.zeros(B, C, H, W, 4)
    for c in range(C):
        rgba = torch.tensor(cmap(c) if colors is None else colors[c % len(colors)])[:4]
        on = (img[:, c] > 0).float()
        masks[:, c] = on[..., None] * rgba
        masks[:, c, ..., 3] = on * alpha
    for b, ax in enumerate(axs):
        for c in range(C):
            ax.imshow(masks[b, c])
    return masks if return_masks else axs
