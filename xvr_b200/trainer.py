"""The training hot loop of xvr on the B200 kernels: host-side mirror of ``Trainer.step`` / ``render_samples``
(/root/reference/src/xvr/model/trainer.py:185-304), ``PoseRegressor`` (model/network.py:7-54),
``PoseRegressionLoss`` / ``DiceLoss`` (model/loss.py:5-89) and ``WarmupCosineSchedule`` (model/scheduler.py:6-30),
sharded over one process per GPU.

What runs where: pose sampling on the host (as the reference), HU -> density, two DRR batches per step through the
CUDA renderer (the first under no_grad, the second with the pose Jacobian), mNCC + Dice + geodesic loss, backward
into the CNN (cuDNN/cuBLAS via PyTorch -- a dense-contraction consumer of the path, not part of it).  Multi-GPU:
poses shard across ranks with a replicated volume; the only collectives are one all-reduce of the flattened CNN
gradients per optimiser step, the global kept-sample count and (optionally) the batch-global min/max that
``Standardize`` needs.  Image augmentations (kornia) and file IO are outside the hot path and not reproduced.
"""

import math

import torch
import torch.distributed as dist

from .data import transform_hu_to_density
from .metrics import DoubleGeodesicSE3, MultiscaleNormalizedCrossCorrelation2d
from .pose import N_ANGULAR_COMPONENTS, convert
from .sampler import random_pose_params
from .sharding import shard_bounds

__all__ = ["PoseRegressor", "PoseRegressionLoss", "DiceLoss", "WarmupCosineSchedule", "adaptive_clip_grad_",
           "render_samples", "TrainStep"]


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class PoseRegressor(torch.nn.Module):
    """CNN backbone (1 input channel, global-average-pooled features) + two linear heads -> RigidTransform."""

    def __init__(self, model_name="resnet18", parameterization="quaternion_adjugate", convention="ZXY", pretrained=False,
                 height=256, unit_conversion_factor=1000.0, norm_layer="groupnorm", **kwargs):
        super().__init__()
        import torchvision  # noqa: PLC0415 - the reference uses timm, which is not available offline

        if pretrained:
            raise NotImplementedError("pretrained backbones need network access")
        if norm_layer == "groupnorm":
            norm = lambda c: torch.nn.GroupNorm(32, c)  # noqa: E731
        elif norm_layer in ("batchnorm", "batchnorm2d", None):
            norm = None
        else:
            raise ValueError(f"unknown norm_layer {norm_layer!r}")
        backbone = getattr(torchvision.models, model_name)(norm_layer=norm, **kwargs)
        old = backbone.conv1
        backbone.conv1 = torch.nn.Conv2d(1, old.out_channels, old.kernel_size, old.stride, old.padding, bias=False)
        features = backbone.fc.in_features
        backbone.fc = torch.nn.Identity()
        self.backbone = backbone
        self.parameterization, self.convention = parameterization, convention
        self.xyz_regression = torch.nn.Linear(features, 3)
        self.rot_regression = torch.nn.Linear(features, N_ANGULAR_COMPONENTS[parameterization])
        self.unit_conversion_factor = unit_conversion_factor

    def forward(self, x):
        x = self.backbone(x)
        rot = self.rot_regression(x)
        xyz = self.unit_conversion_factor * self.xyz_regression(x)
        return convert(rot, xyz, parameterization=self.parameterization, convention=self.convention)


class DiceLoss(torch.nn.Module):
    """1 - mean foreground Dice between two multi-channel masks (channel 0 = background)."""

    def forward(self, a, b):
        a, b = a.flatten(2).to(torch.float32), b.flatten(2).to(torch.float32)
        dice = (2.0 * (a * b).sum(2) / (a.sum(2) + b.sum(2)))[:, 1:]
        return 1 - dice.nanmean(dim=1).nan_to_num()


class PoseRegressionLoss(torch.nn.Module):
    def __init__(self, sdd, weight_ncc=1.0, weight_geo=1e-2, weight_dice=1.0, weight_mvc=0.0):
        super().__init__()
        self.imagesim = MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])
        self.diceloss = DiceLoss()
        self.geodesic = DoubleGeodesicSE3(sdd)
        self.weight_ncc, self.weight_geo, self.weight_dice, self.weight_mvc = weight_ncc, weight_geo, weight_dice, weight_mvc

    def forward(self, img, mask, pose, pred_img, pred_mask, pred_pose):
        mncc = self.imagesim(img, pred_img)
        dice = self.diceloss(mask, pred_mask)
        rgeo, tgeo, dgeo = self.geodesic(pose, pred_pose)
        loss = self.weight_ncc * (1 - mncc) + self.weight_dice * dice + self.weight_geo * dgeo
        mvc = self.multiview_consistency(pose, pred_pose)
        if self.weight_mvc > 0:
            loss = loss + self.weight_mvc * mvc.mean()
        return loss, mncc, dgeo, rgeo, tgeo, dice, mvc

    def multiview_consistency(self, true_pose, pred_pose):
        B = len(true_pose)
        if B < 2:
            return torch.zeros(1, device=true_pose.matrix.device)
        i, j = torch.triu_indices(B, B, offset=1)
        return self.geodesic(true_pose[j] @ true_pose[i].inverse(), pred_pose[j] @ pred_pose[i].inverse())[2]


class WarmupCosineSchedule(torch.optim.lr_scheduler.LambdaLR):
    """Linear warm-up to the base rate over ``warmup_steps`` then a half-cosine decay to 0 at ``t_total``."""

    def __init__(self, optimizer, warmup_steps, t_total, cycles=0.5, last_epoch=-1):
        self.warmup_steps, self.t_total, self.cycles = warmup_steps, t_total, cycles
        super().__init__(optimizer, self._factor, last_epoch=last_epoch)

    def _factor(self, step):
        if step < self.warmup_steps:
            return step / max(1.0, self.warmup_steps)
        progress = (step - self.warmup_steps) / max(1, self.t_total - self.warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 2.0 * self.cycles * progress)))


def _unitwise_norm(x):
    if x.ndim <= 1:
        return x.norm(2.0)
    return x.norm(2.0, dim=tuple(range(1, x.ndim)), keepdim=True)


@torch.no_grad()
def adaptive_clip_grad_(parameters, clip_factor=0.01, eps=1e-3):
    """Adaptive gradient clipping (unit-wise ||g|| <= clip_factor * max(||w||, eps)); the reference calls timm's
    ``adaptive_clip_grad_`` at trainer.py:227."""
    for p in parameters:
        if p.grad is None:
            continue
        max_norm = _unitwise_norm(p.detach()).clamp_(min=eps).mul_(clip_factor)
        g_norm = _unitwise_norm(p.grad)
        clipped = p.grad * (max_norm / g_norm.clamp(min=1e-6))
        p.grad.copy_(torch.where(g_norm < max_norm, p.grad, clipped))


def render_samples(drr, volume, seg, affinv, pose, img_threshold=0.10, mask_threshold=0.05):
    """trainer.py:279-304 verbatim in behaviour: rays -> renderer -> (img (B,1,H,W), mask (B,C,H,W) bool, keep (B,))."""
    source, target = drr.detector(pose, None)
    raylen = (target - source).norm(dim=-1).unsqueeze(1)
    source, target = affinv(source), affinv(target)
    img = drr.renderer(volume, source, target, raylen, mask=seg)
    img = drr.reshape_transform(img, batch_size=len(pose))
    mask = img > 0
    img = img.sum(dim=1, keepdim=True)
    if mask.shape[1] == 1:
        keep = mask.to(img).flatten(1).mean(1) > img_threshold
    else:
        keep = (mask[:, 1:].sum(dim=1, keepdim=True) > 0).to(img).flatten(1).mean(1) > mask_threshold
    return img, mask, keep


class TrainStep:
    """One rank's share of the xvr training iteration (``Trainer.step``), gradient accumulation included.

    ``volumes`` is a list of (hu (D,H,W) CUDA tensor, labelmap or None, affine_inverse RigidTransform, offset
    RigidTransform) -- what ``Trainer.load`` (trainer.py:248-277) produces per subject.  Every rank must pass the
    same list; subject choice and contrast are drawn from a generator seeded identically on all ranks, the pose
    batch from a per-rank generator.
    """

    def __init__(self, drr, model, volumes, pose_distribution, transforms, sdd, batch_size=116, lr=2e-4,
                 n_total_itrs=1_000_000, n_warmup_itrs=1_000, n_grad_accum_itrs=4, weight_ncc=1.0, weight_geo=1e-2,
                 weight_dice=1.0, weight_mvc=0.0, seed=0, standardize_global=True, disable_scheduler=False):
        self.rank, self.world = _world()
        self.drr, self.model, self.volumes, self.transforms = drr, model, volumes, transforms
        self.pose_distribution = dict(pose_distribution)
        self.batch_size = batch_size
        lo, hi = shard_bounds(batch_size, self.rank, self.world)
        self.local_batch = hi - lo
        self.lossfn = PoseRegressionLoss(sdd, weight_ncc, weight_geo, weight_dice, weight_mvc)
        self.optimizer = torch.optim.Adam(model.parameters(), lr=lr)
        if disable_scheduler:
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lambda step: 1.0)
        else:
            self.scheduler = WarmupCosineSchedule(self.optimizer, n_warmup_itrs / n_grad_accum_itrs,
                                                  n_total_itrs / n_grad_accum_itrs)
        self.n_total_itrs, self.n_grad_accum_itrs = n_total_itrs, n_grad_accum_itrs
        self.shared_rng = torch.Generator().manual_seed(seed)           # same stream on every rank
        self.pose_rng = torch.Generator().manual_seed(seed * 9973 + 1 + self.rank)
        self.standardize_global = standardize_global and self.world > 1
        self.device = next(model.parameters()).device

    # ------------------------------------------------------------------ pieces
    def _standardize(self, x):
        """XrayTransforms with the batch-global min/max of the UNSHARDED batch (utils/preprocess.py:28-29)."""
        if not self.standardize_global:
            return self.transforms(x)
        lo, hi = x.detach().min(), x.detach().max()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        x = (x - lo) / (hi - lo + 1e-6)
        t = self.transforms
        if tuple(x.shape[-2:]) != t.size:
            x = torch.nn.functional.interpolate(x, size=t.size, mode="bilinear", align_corners=False, antialias=True)
        return (x - t.mean) / t.std

    def _allreduce_grads(self):
        if self.world == 1:
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)  # NCCL over NVLink; ~45 MB for ResNet-18
        offset = 0
        for g in grads:
            g.copy_(flat[offset:offset + g.numel()].view_as(g))
            offset += g.numel()

    # ------------------------------------------------------------------ the iteration
    def step(self, itr):
        dev = self.device
        subject = int(torch.randint(len(self.volumes), (1,), generator=self.shared_rng))
        contrast = float(torch.empty(1).uniform_(1.0, 10.0, generator=self.shared_rng))
        vol, seg, affinv, offset = self.volumes[subject]

        rot, xyz = random_pose_params(**self.pose_distribution, batch_size=self.local_batch, generator=self.pose_rng)
        pose = convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY", degrees=True)
        pose = pose.compose(offset)

        density = transform_hu_to_density(vol, contrast)
        with torch.no_grad():
            img, mask, keep = render_samples(self.drr, density, seg, affinv, pose)
        img, mask, pose = img[keep], mask[keep], pose[keep]

        kept = torch.tensor([float(keep.sum())], device=dev)
        if self.world > 1:
            dist.all_reduce(kept, op=dist.ReduceOp.SUM)
        log = {"kept": kept.item() / self.batch_size}
        if len(pose) > 0:
            x = self._standardize(img)
            pred_pose = self.model(x)
            pred_img, pred_mask, _ = render_samples(self.drr, density, seg, affinv, pred_pose)
            x_true, x_pred = x, self._standardize(pred_img)
            loss, mncc, dgeo, rgeo, tgeo, dice, mvc = self.lossfn(x_true, mask, pose, x_pred, pred_mask, pred_pose)
            # global mean over the kept samples of ALL ranks, scaled for gradient accumulation
            (loss.sum() / kept.clamp_min(1.0) / self.n_grad_accum_itrs).backward()
            sums = torch.stack([loss.sum(), mncc.sum(), dgeo.sum(), rgeo.sum(), tgeo.sum(), dice.sum()]).detach()
        else:
            sums = torch.zeros(6, device=dev)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        means = (sums / kept.clamp_min(1.0)).tolist()
        log.update(dict(zip(("loss", "mncc", "dgeo", "rgeo", "tgeo", "dice"), means)))

        if (itr + 1) % self.n_grad_accum_itrs == 0 or (itr + 1) == self.n_total_itrs:
            self._allreduce_grads()
            adaptive_clip_grad_(self.model.parameters())
            self.optimizer.step()
            self.scheduler.step()
            self.optimizer.zero_grad()
        log["lr"] = self.scheduler.get_last_lr()[0]
        return log
