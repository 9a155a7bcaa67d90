"""Oracle restatement of the DiffDRR render path on real ``grid_sample`` -- TEST INFRASTRUCTURE ONLY.

Follows diffdrr 0.6.0 ``diffdrr/detector.py`` (Detector._initialize_carm / forward),
``diffdrr/renderers.py`` (Siddon, Trilinear, _get_alphas, _get_alpha_minmax, _get_xyzs, _get_voxel),
``diffdrr/drr.py`` (DRR.forward / render / reshape_transform) and ``diffdrr/data.py``
(transform_hu_to_density), in the exact call sequence xvr re-states at
/root/reference/src/xvr/model/trainer.py:279-304 (detector -> ray length -> affine_inverse ->
renderer -> reshape_transform).  PARITY UNPINNED (see oracle/__init__.py); every item recalled
from memory is a knob in oracle/knobs.py.
"""

import torch
import torch.nn.functional as F

from . import knobs
from .pose import apply, compose

__all__ = [
    "REORIENT",
    "detector_grid",
    "detector_rays",
    "alpha_minmax",
    "trilinear_render",
    "siddon_alphas",
    "siddon_render",
    "siddon_segments",
    "drr_forward",
    "hu_to_density",
    "standardize",
    "xray_transforms",
]

# diffdrr/data.py read(orientation=...)  
REORIENT = {
    "AP": torch.tensor([[1.0, 0, 0, 0], [0, 0, -1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]]),
    "PA": torch.tensor([[1.0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]]),
    None: torch.eye(4),
}


# ----------------------------------------------------------------------------- detector (a2)
def detector_grid(height, width, delx, dely, x0, y0, sdd, reverse_x_axis, dtype=torch.float32):
    """Calibrated detector plane in the camera frame, (1, H*W, 3), row-major (i rows, j cols).

    diffdrr/detector.py _initialize_carm + calibration matrix [[delx,0,0,x0],[0,dely,0,y0],[0,0,sdd,0]]
    applied to (s_j, t_i, 1).
    """
    h_off = 1.0 if height % 2 else 0.5
    w_off = 1.0 if width % 2 else 0.5
    t = torch.arange(-height // 2, height // 2, dtype=dtype) + h_off
    s = torch.arange(-width // 2, width // 2, dtype=dtype) + w_off
    t = knobs.DET_SIGN_T * t
    s = knobs.DET_SIGN_S * s
    if reverse_x_axis:
        s = -s
    coefs = torch.cartesian_prod(t, s).reshape(-1, 2)  # (t_i, s_j), i major
    target = torch.stack(
        [coefs[:, 1] * delx + x0, coefs[:, 0] * dely + y0, torch.full_like(coefs[:, 0], 1.0) * sdd], -1
    )
    return target[None]


def detector_rays(pose, reorient, height, width, delx, dely, x0, y0, sdd, reverse_x_axis):
    """Detector.forward(pose, None) -> source (B,1,3), target (B,H*W,3) in world mm."""
    tgt = detector_grid(height, width, delx, dely, x0, y0, sdd, reverse_x_axis, pose.dtype).to(pose.device)
    src = torch.zeros(1, 1, 3, dtype=pose.dtype, device=pose.device)
    full = compose(reorient.to(pose)[None], pose)  # reorient.compose(extrinsic)
    B = full.shape[0]
    return apply(full, src.expand(B, -1, -1)), apply(full, tgt.expand(B, -1, -1))


# ----------------------------------------------------------------------------- shared helpers
def alpha_minmax(source, target, dims, eps):
    """_get_alpha_minmax: slab test of the segment against the box [0, dims], clamped to [0,1]."""
    sdd = target - source + eps
    alpha0 = (torch.zeros(3).to(source) - source) / sdd
    alpha1 = (dims.to(source) - source) / sdd
    alphas = torch.stack([alpha0, alpha1])
    alphamin = alphas.min(dim=0).values.max(dim=-1).values.unsqueeze(-1)
    alphamax = alphas.max(dim=0).values.min(dim=-1).values.unsqueeze(-1)
    alphamin = torch.where(alphamin < 0.0, torch.zeros_like(alphamin), alphamin)
    alphamax = torch.where(alphamax > 1.0, torch.ones_like(alphamax), alphamax)
    return alphamin, alphamax


def _xyzs(alpha, source, target, eps):
    """Points along the ray: s + alpha * (t - s + eps), (B,N,n,3) in voxel-index coordinates."""
    return source.unsqueeze(-2) + alpha.unsqueeze(-1) * (target - source + eps).unsqueeze(2)


def _voxel(volume, grid, mode, align_corners):
    """_get_voxel: grid_sample on volume.permute(2,1,0) so grid[...,0] indexes volume dim 0."""
    B = grid.shape[0]
    out = F.grid_sample(
        volume.permute(2, 1, 0)[None, None].expand(B, -1, -1, -1, -1),
        grid.unsqueeze(1),
        mode=mode,
        padding_mode="zeros",
        align_corners=align_corners,
    )
    return out[:, 0, 0]


def _to_channels(samples, volume, mask, grid, align_corners):
    """A8: scatter the weighted samples into label channels; C = mask.max()+1."""
    B, N, _ = samples.shape
    C = int(mask.max().item() + 1)
    ch = _voxel(mask.to(volume), grid, "nearest", align_corners).long()
    out = torch.zeros(B, C, N).to(samples)
    return out.scatter_add_(1, ch.transpose(-1, -2), samples.transpose(-1, -2))


# ----------------------------------------------------------------------------- trilinear (a4)
def trilinear_render(volume, source, target, raylen, n_points=None, mask=None, step=None, eps=None):
    """Trilinear.forward(volume, source, target, img, n_points=500, align_corners=True, mask=None).

    volume (D0,D1,D2); source (B,1,3) / target (B,N,3) in voxel-index coordinates; raylen (B,1,N)
    is the world-mm ray length xvr computes at trainer.py:284.  Returns (B,C,N).
    """
    n = knobs.TRILINEAR_N_POINTS if n_points is None else n_points
    step = knobs.TRILINEAR_STEP if step is None else step
    eps = knobs.RENDER_EPS if eps is None else eps
    dims = torch.tensor(volume.shape).to(volume) - 1
    amin, amax = alpha_minmax(source, target, dims, eps)
    alphas = torch.linspace(0, 1, n)[None, None].to(volume) * (amax - amin) + amin
    xyzs = _xyzs(alphas, source, target, eps)
    grid = 2 * xyzs / dims - 1
    samples = _voxel(volume, grid, "bilinear", knobs.TRILINEAR_ALIGN_CORNERS)  # (B,N,n)
    if mask is None:
        img = samples.sum(dim=-1).unsqueeze(1)
    else:
        img = _to_channels(samples, volume, mask, grid, knobs.TRILINEAR_ALIGN_CORNERS)
    span = (amax - amin)[..., 0].unsqueeze(1)  # (B,1,N)
    if step == "span/(n-1)":
        w = span / (n - 1)
    elif step == "span/n":
        w = span / n
    elif step == "1/n":
        w = torch.full_like(span, 1.0 / n)
    else:
        raise ValueError(step)
    return img * raylen * w


# ----------------------------------------------------------------------------- Siddon (a5)
def siddon_alphas(source, target, shape, voxel_shift, eps):
    """_get_alphas: every plane crossing, out-of-range -> NaN, sorted, all-NaN columns dropped."""
    pieces = []
    for a in range(3):
        planes = torch.arange(shape[a] + 1).to(source) - voxel_shift
        sa, ta = source[..., a : a + 1], target[..., a : a + 1]
        pieces.append((planes.expand(len(source), 1, -1) - sa) / (ta - sa + eps))
    alphas = torch.cat(pieces, dim=-1)
    lo = torch.zeros(3).to(source) - voxel_shift
    hi = torch.tensor(shape).to(source) - voxel_shift
    sdd = target - source + eps
    a0, a1 = (lo - source) / sdd, (hi - source) / sdd
    st = torch.stack([a0, a1])
    amin = st.min(dim=0).values.max(dim=-1).values.unsqueeze(-1)
    amax = st.max(dim=0).values.min(dim=-1).values.unsqueeze(-1)
    amin = torch.where(amin < 0.0, torch.zeros_like(amin), amin)
    amax = torch.where(amax > 1.0, torch.ones_like(amax), amax)
    good = torch.logical_and(alphas >= amin, alphas <= amax)
    alphas = torch.where(good, alphas, torch.full_like(alphas, float("nan")))
    alphas = torch.sort(alphas, dim=-1).values  # NaNs sort last
    keep = ~alphas.isnan().all(dim=0).all(dim=0)
    return alphas[..., keep]


def _siddon_grid(alphamid, source, target, shape, voxel_shift, eps):
    xyzs = _xyzs(alphamid, source, target, eps)
    dims = torch.tensor(shape).to(source)
    return 2 * (xyzs + voxel_shift) / dims - 1


def siddon_render(volume, source, target, raylen, mask=None, voxel_shift=None, eps=None):
    """Siddon.forward(volume, source, target, img, align_corners=False, mask=None) -> (B,C,N)."""
    vs = knobs.SIDDON_VOXEL_SHIFT_DEFAULT if voxel_shift is None else voxel_shift
    eps = knobs.RENDER_EPS if eps is None else eps
    alphas = siddon_alphas(source, target, tuple(volume.shape), vs, eps)
    alphamid = (alphas[..., 0:-1] + alphas[..., 1:]) / 2
    grid = _siddon_grid(alphamid, source, target, tuple(volume.shape), vs, eps)
    # NaN midpoints -> out of bounds -> zero padding.  Nearest-neighbour lookups have a zero gradient w.r.t. the
    # grid; detaching states that directly (autograd would otherwise form 0 * NaN = NaN on the padded columns).
    grid = torch.nan_to_num(grid, nan=-2.0).detach()
    voxels = _voxel(volume, grid, "nearest", knobs.SIDDON_ALIGN_CORNERS)
    seg = torch.diff(alphas, dim=-1)
    weighted = torch.nan_to_num(voxels * seg, nan=0.0)
    if mask is None:
        img = weighted.sum(dim=-1).unsqueeze(1)
    else:
        img = _to_channels(weighted, volume, mask, grid, knobs.SIDDON_ALIGN_CORNERS)
    return img * raylen


def siddon_segments(volume_shape, source, target, voxel_shift=None, eps=None):
    """Per-ray (flat voxel index, segment length in alpha) lists -- the bit-exactness reference.

    Reconstructs the index grid_sample(nearest, align_corners=False) looks up
    (ATen/native/cuda/GridSampler.cuh:23-31: ((g+1)*size-1)/2 then nearbyint) for every segment.
    Returns idx (B,N,M) int64 with -1 where out of bounds / NaN, and seg (B,N,M) fp32 (0 where NaN).
    """
    vs = knobs.SIDDON_VOXEL_SHIFT_DEFAULT if voxel_shift is None else voxel_shift
    eps = knobs.RENDER_EPS if eps is None else eps
    alphas = siddon_alphas(source, target, tuple(volume_shape), vs, eps)
    alphamid = (alphas[..., 0:-1] + alphas[..., 1:]) / 2
    g = _siddon_grid(alphamid, source, target, tuple(volume_shape), vs, eps)
    size = torch.tensor(volume_shape).to(source)
    u = ((g + 1) * size - 1) / 2
    r = torch.round(u)  # ties-to-even == nearbyint
    inb = ((r >= 0) & (r <= size - 1)).all(-1) & ~alphamid.isnan()
    r = torch.nan_to_num(r, nan=0.0).long()
    flat = (r[..., 0] * volume_shape[1] + r[..., 1]) * volume_shape[2] + r[..., 2]
    flat = torch.where(inb, flat, torch.full_like(flat, -1))
    seg = torch.nan_to_num(torch.diff(alphas, dim=-1), nan=0.0)
    return flat, seg


# ----------------------------------------------------------------------------- DRR (a2-a7)
def drr_forward(
    volume,
    affine_inverse,
    pose,
    *,
    reorient,
    height,
    width,
    delx,
    dely,
    x0,
    y0,
    sdd,
    reverse_x_axis,
    renderer="trilinear",
    mask=None,
    **kw,
):
    """DRR.forward(pose) following trainer.py:283-289; returns (B,C,H,W)."""
    source, target = detector_rays(pose, reorient, height, width, delx, dely, x0, y0, sdd, reverse_x_axis)
    raylen = (target - source).norm(dim=-1).unsqueeze(1)
    source = apply(affine_inverse, source)
    target = apply(affine_inverse, target)
    if renderer == "trilinear":
        img = trilinear_render(volume, source, target, raylen, mask=mask, **kw)
    elif renderer == "siddon":
        img = siddon_render(volume, source, target, raylen, mask=mask, **kw)
    else:
        raise ValueError(renderer)
    return img.view(pose.shape[0], -1, height, width)


# ----------------------------------------------------------------------------- a8
def hu_to_density(volume, bone_attenuation_multiplier):
    """diffdrr/data.py transform_hu_to_density (called every step at trainer.py:196-197)."""
    volume = volume.to(torch.float32)
    air = torch.where(volume <= knobs.HU_AIR)
    soft = torch.where((knobs.HU_AIR < volume) & (volume <= knobs.HU_BONE))
    bone = torch.where(knobs.HU_BONE < volume)
    density = torch.empty_like(volume)
    density[air] = volume[soft].min()
    density[soft] = volume[soft]
    density[bone] = volume[bone] * bone_attenuation_multiplier
    density -= density.min()
    density /= density.max()
    return density


# ----------------------------------------------------------------------------- a12 (xvr-owned)
def standardize(x, eps=1e-6):
    """/root/reference/src/xvr/utils/preprocess.py:23-29 (batch-global min/max)."""
    return (x - x.min()) / (x.max() - x.min() + eps)


def xray_transforms(x, height, width=None, mean=0.15, std=0.1):
    """XrayTransforms without Equalize: Standardize -> Resize -> Normalize (preprocess.py:5-20)."""
    width = height if width is None else width
    x = standardize(x)
    if x.shape[-2:] != (height, width):
        x = F.interpolate(x, size=(height, width), mode="bilinear", align_corners=False, antialias=True)
    return (x - mean) / std
