"""Volume containers and intensity mapping -- the slice of ``diffdrr.data`` that sits on the hot path.

``transform_hu_to_density`` is called every training step (/root/reference/src/xvr/model/trainer.py:196-197).
``Subject``/``Volume`` carry what ``DRR`` and xvr read from a torchio subject
(/root/reference/src/xvr/model/utils.py:162-171: ``subject.volume.data``, ``subject.volume.get_center()``).
File IO (NIfTI via torchio) is outside the hot path; ``read`` accepts tensors, and paths only when torchio
is importable.
"""

import weakref

import numpy as np
import torch

from . import _conventions as conv

__all__ = ["Volume", "Subject", "read", "load_example_ct", "transform_hu_to_density", "synthetic_ct", "REORIENT"]

REORIENT = {
    "AP": [[1.0, 0, 0, 0], [0, 0, -1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]],
    "PA": [[1.0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]],
    None: [[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]],
}


class Volume:
    """A (1,D0,D1,D2) image with a voxel-index -> world-mm affine (torchio.ScalarImage look-alike)."""

    def __init__(self, data, affine):
        if data.dim() == 3:
            data = data[None]
        self.data = data
        self.affine = np.asarray(affine, dtype=np.float64)

    @property
    def spatial_shape(self):
        return tuple(self.data.shape[1:])

    def get_center(self):
        """World coordinates of the volume centre (voxel (shape-1)/2)."""
        c = (np.array(self.spatial_shape, dtype=np.float64) - 1) / 2
        return tuple((self.affine[:3, :3] @ c + self.affine[:3, 3]).tolist())


class Subject:
    def __init__(self, volume, mask=None, density=None, reorient=None, orientation="AP", fiducials=None):
        self.volume = volume
        self.mask = mask
        self.density = density
        self.reorient = reorient
        self.orientation = orientation
        self.fiducials = fiducials


class _HuStats:
    """{min soft, max soft, min bone, max bone} per HU volume (they do not depend on the multiplier, and training
    calls transform_hu_to_density on the same few volumes every step): one reduction per volume object and
    version, kept while the volume is alive."""

    def __init__(self):
        self.cache = {}  # id(volume) -> (weakref to the volume, version, stats); entries die with their volume

    def get(self, volume):
        key = id(volume)
        hit = self.cache.get(key)
        if hit is None or hit[0]() is not volume or hit[1] != volume._version:
            from ._lib import call, ptr, stream  # noqa: PLC0415

            stats = torch.empty(4, device=volume.device, dtype=torch.float32)
            work = torch.empty(4, device=volume.device, dtype=torch.int32)
            call("xvr_hu_stats", ptr(volume), volume.numel(), conv.HU_AIR, conv.HU_BONE, ptr(work), ptr(stats), stream())
            hit = (weakref.ref(volume, lambda _, k=key, c=self.cache: c.pop(k, None)), volume._version, stats)
            self.cache[key] = hit
        return hit[2]


_hu_stats = _HuStats()


def transform_hu_to_density(volume, bone_attenuation_multiplier, out=None):
    """Piecewise HU -> density map, shifted and scaled to [0,1].

    air (HU <= -800) takes the minimum soft-tissue value, soft tissue (-800, 350] is kept, bone (> 350) is
    multiplied by ``bone_attenuation_multiplier`` (a float, or a 1-element CUDA tensor that the kernel reads at run
    time -- what a CUDA-graph-captured training step needs).  ``out`` optionally receives the result.  CUDA volumes (the per-step call of the training loop) go
    through the one-pass kernel of csrc/density.cu; CPU volumes (``read`` at set-up time, before ``.to(device)``)
    are mapped with tensor ops.
    """
    if volume.is_cuda:
        from ._lib import call, ptr, stream  # noqa: PLC0415

        vol = volume if (volume.dtype == torch.float32 and volume.is_contiguous()) else volume.float().contiguous()
        if out is None:
            out = torch.empty_like(vol)
        elif out.shape != vol.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != vol.device:
            raise ValueError("out must be a contiguous float32 tensor of the volume's shape on the volume's device")
        m_dev = None
        if torch.is_tensor(bone_attenuation_multiplier):
            m_dev = bone_attenuation_multiplier
            if not m_dev.is_cuda or m_dev.dtype != torch.float32 or m_dev.numel() != 1:
                raise ValueError("a tensor multiplier must be a 1-element float32 CUDA tensor")
        call("xvr_hu_to_density", ptr(vol), vol.numel(), conv.HU_AIR, conv.HU_BONE,
             0.0 if m_dev is not None else float(bone_attenuation_multiplier), ptr(m_dev), ptr(_hu_stats.get(vol)),
             ptr(out), stream())
        return out
    volume = volume.to(torch.float32)
    soft = (volume > conv.HU_AIR) & (volume <= conv.HU_BONE)
    bone = volume > conv.HU_BONE
    soft_min = torch.where(soft, volume, torch.full_like(volume, float("inf"))).min()
    density = torch.where(bone, volume * bone_attenuation_multiplier, torch.where(soft, volume, soft_min))
    density = density - density.min()
    return density / density.max()


def read(volume, labelmap=None, labels=None, orientation="AP", bone_attenuation_multiplier=1.0, fiducials=None,
         affine=None, center_volume=True, **kwargs):
    """Build a :class:`Subject` from tensors (or from image paths when torchio is importable).

    Mirrors ``diffdrr.data.read(volume, mask, labels, orientation, **kw)`` as used at
    /root/reference/src/xvr/renderer/load.py:26: canonical orientation, the isocenter moved to the world
    origin, ``density = transform_hu_to_density(volume, m)``, and ``labels`` restricting the density to the
    selected structures.
    """
    if isinstance(volume, (str, bytes)) or hasattr(volume, "__fspath__"):
        try:
            import torchio  # noqa: PLC0415
        except ImportError as e:  # pragma: no cover - torchio is not in the build image
            raise ImportError("reading image files needs torchio; pass tensors + affine instead") from e
        img = torchio.ToCanonical()(torchio.ScalarImage(volume))
        data, affine = img.data, img.affine
        if labelmap is not None:
            labelmap = torchio.ToCanonical()(torchio.LabelMap(labelmap)).data
    elif isinstance(volume, Volume):
        data, affine = volume.data, volume.affine
    else:
        data = volume
        if affine is None:
            affine = np.eye(4)
    affine = np.array(affine, dtype=np.float64)
    vol = Volume(data, affine)
    if center_volume:
        affine = affine.copy()
        affine[:3, 3] -= np.array(vol.get_center())
        vol = Volume(data, affine)
    density = transform_hu_to_density(vol.data[0], bone_attenuation_multiplier)
    mask = None
    if labelmap is not None:
        mask_data = labelmap if torch.is_tensor(labelmap) else labelmap.data
        mask = Volume(mask_data if mask_data.dim() == 4 else mask_data[None], affine)
        if labels is not None:
            if isinstance(labels, int):
                labels = [labels]
            keep = torch.isin(mask.data[0], torch.as_tensor(labels).to(mask.data))
            density = density * keep
    if orientation not in REORIENT:
        raise ValueError(f"Unrecognized orientation {orientation!r}")
    return Subject(vol, mask=mask, density=density, reorient=torch.tensor(REORIENT[orientation]),
                   orientation=orientation, fiducials=fiducials)


def load_example_ct(*args, **kwargs):
    """``diffdrr.data.load_example_ct`` downloads nothing here: the example CT ships with the DiffDRR wheel, which is
    not part of this package (xvr only uses it as a placeholder subject for multi-subject training,
    /root/reference/src/xvr/model/utils.py:155).  Use ``read(...)`` on your own volume or ``synthetic_ct``."""
    raise RuntimeError("xvr_b200 does not bundle DiffDRR's example CT; pass a volume to read() or use synthetic_ct()")


def synthetic_ct(n, seed=0, with_labels=False, device="cpu"):
    """HU-like phantom on an n^3 grid spanning 256 mm: air, a soft-tissue ellipsoid with
    noise and six bone ellipsoids.  Returns (hu (n,n,n), labelmap or None, affine (4,4) numpy)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ax = torch.linspace(-1, 1, n, device=device)
    X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    hu = torch.full((n, n, n), -1000.0, device=device)
    lab = torch.zeros((n, n, n), dtype=torch.uint8, device=device) if with_labels else None
    soft = (X / 0.9) ** 2 + (Y / 0.7) ** 2 + (Z / 0.9) ** 2 < 1
    noise = (torch.randn((n, n, n), generator=g) * 20.0).to(device)
    hu = torch.where(soft, 40.0 + noise, hu)
    del noise
    if lab is not None:
        lab[soft] = 1
    for k in range(6):
        c = ((torch.rand(3, generator=g) - 0.5) * 0.9).tolist()
        r = (0.08 + 0.12 * torch.rand(3, generator=g)).tolist()
        val = 700.0 + 800.0 * torch.rand(1, generator=g).item()
        bone = ((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2 < 1
        hu = torch.where(bone, torch.full_like(hu, val), hu)
        if lab is not None:
            lab[bone] = 2 + k
    sp = 256.0 / n
    affine = np.diag([sp, sp, sp, 1.0])
    return hu, lab, affine
