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
from .renderers import Trilinear
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
                 height=256, unit_conversion_factor=1000.0, norm_layer="groupnorm", channels_last=False,
                 backbone=None, autocast_bf16=False, **kwargs):
        super().__init__()
        if backbone is not None:
            # any feature extractor (B,1,H,W) -> (B,F); F probed like the reference does (network.py:40)
            with torch.no_grad():
                features = backbone(torch.randn(1, 1, height, height)).shape[-1]
        else:
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
        # NHWC activations: cuDNN's tensor-core convolutions stop converting NCHW <-> NHWC around every layer
        self.channels_last = bool(channels_last)
        if self.channels_last:
            self.backbone.to(memory_format=torch.channels_last)
        # bf16 autocast around the BACKBONE only (dense convolutions: the one place tensor cores belong on this path);
        # the two regression heads and the pose conversion stay in fp32.  Off by default: the reference trains in fp32.
        self.autocast_bf16 = bool(autocast_bf16)

    def forward(self, x):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        if self.autocast_bf16 and x.is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x = self.backbone(x)
            x = x.float()
        else:
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
        if self.weight_mvc > 0:
            mvc = self.multiview_consistency(pose, pred_pose)
            loss = loss + self.weight_mvc * mvc.mean()
        else:  # the reference evaluates it for its log only (B(B-1)/2 pose pairs); skipped when it carries no weight
            mvc = loss.new_zeros(1)
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


class _RenderEpilogue(torch.autograd.Function):
    """(B,C,N) rendered channels -> (channel sum (B,1,N), stats (B,4) = {foreground fraction, keep, min, max}) in one
    launch (csrc/train.cu); the backward hands the gradient of the sum back as an expanded view, which the renderer's
    backward recognises (one upstream gradient per ray -> the saved Jacobian serves)."""

    @staticmethod
    def forward(ctx, img, img_threshold, mask_threshold):
        from ._lib import call, cuda_f32, ptr, stream  # noqa: PLC0415

        img = cuda_f32(img, "rendered batch")
        B, C, N = img.shape
        ctx.C = C
        stats = torch.empty(B, 4, device=img.device, dtype=torch.float32)
        total = img if C == 1 else torch.empty(B, 1, N, device=img.device, dtype=torch.float32)
        if B > 0:
            call("xvr_render_epilogue", ptr(img), B, C, N, float(img_threshold), float(mask_threshold),
                 None if C == 1 else ptr(total), ptr(stats), stream())
        ctx.mark_non_differentiable(stats)
        return total.view(B, 1, N), stats

    @staticmethod
    def backward(ctx, gsum, _gstats):
        return gsum.expand(-1, ctx.C, -1), None, None


def render_samples_fused(drr, volume, seg, affinv, pose, img_threshold=0.10, mask_threshold=0.05):
    """``render_samples`` through the fused kernels: the rays are generated inside the renderer (no (B,N,3) tensors:
    0.5 GB per batch at 256^2; label channels too, for the trilinear renderer), and mask / channel sum / keep / per-sample min and max come out of
    one epilogue launch.  Returns ``(img (B,1,H,W), mask (B,C,H,W) bool or None, keep (B,) bool, stats (B,4))``;
    ``mask`` is only formed with label channels (the Dice term is a constant without them)."""
    B = len(pose)
    fused_labels = seg is not None and isinstance(drr.renderer, Trilinear) and not volume.requires_grad
    if (seg is None or fused_labels) and hasattr(drr.renderer, "render_drr"):
        cam2world = drr.detector.reorient.compose(pose).matrix
        cam2vox = affinv.matrix.to(cam2world) @ cam2world
        raw = drr.renderer.render_drr(volume, cam2vox[:, :3].contiguous(), cam2world[:, :3].contiguous(), drr.detector,
                                      **({"mask": seg} if fused_labels else {}))
    else:
        source, target = drr.detector(pose, None)
        raylen = (target - source).norm(dim=-1).unsqueeze(1)
        raw = drr.renderer(volume, affinv(source), affinv(target), raylen, mask=seg)
    total, stats = _RenderEpilogue.apply(raw, img_threshold, mask_threshold)
    H, W = drr.detector.height, drr.detector.width
    mask = (raw.detach() > 0).view(B, -1, H, W) if seg is not None else None
    return total.view(B, 1, H, W), mask, stats[:, 1] > 0, stats


class TrainStep:
    """One rank's share of the xvr training iteration (``Trainer.step``), gradient accumulation included.

    ``volumes`` is a list of (hu (D,H,W) CUDA tensor, labelmap or None, affine_inverse RigidTransform, offset
    RigidTransform) -- what ``Trainer.load`` (trainer.py:248-277) produces per subject.  Every rank must pass the
    same list; subject choice and contrast are drawn from a generator seeded identically on all ranks, the pose
    batch from a per-rank generator.

    ``shard`` chooses how several ranks divide the work.  ``"batch"``: every iteration's batch is split over all
    ranks (the renderer scales with it, the CNN at a per-rank batch of ~15 does not).  ``"accumulation"``: the
    ``n_grad_accum_itrs`` iterations between two optimiser steps are independent given the weights, so they are
    dealt out to groups of ranks -- with W <= A ranks every rank runs whole iterations at the full batch, with
    W = g * A each iteration is split over a sub-group of g ranks -- and the gradients meet in ONE all-reduce per
    optimiser step.  Same arithmetic as the reference's gradient accumulation (trainer.py:219-231), a CNN batch g
    times larger than batch sharding gives it.
    """

    def __init__(self, drr, model, volumes, pose_distribution, transforms, sdd, batch_size=116, lr=2e-4,
                 n_total_itrs=1_000_000, n_warmup_itrs=1_000, n_grad_accum_itrs=4, weight_ncc=1.0, weight_geo=1e-2,
                 weight_dice=1.0, weight_mvc=0.0, seed=0, standardize_global=True, disable_scheduler=False,
                 use_cuda_graph=False, log_every=1, shard="batch"):
        self.rank, self.world = _world()
        self.drr, self.model, self.volumes, self.transforms = drr, model, volumes, transforms
        self.pose_distribution = dict(pose_distribution)
        self.batch_size = batch_size
        self.seed = seed
        # ---- who works on what: n_groups groups of sub_world ranks; group g owns the iterations with
        # (itr mod n_grad_accum_itrs) mod n_groups == g and splits their batch over its sub_world ranks
        if shard not in ("batch", "accumulation"):
            raise ValueError("shard must be 'batch' or 'accumulation'")
        self.n_groups, self.sub_world, self.sub_group = 1, self.world, None
        if shard == "accumulation" and self.world > 1:
            A = n_grad_accum_itrs
            if self.world <= A and A % self.world == 0:
                self.n_groups, self.sub_world = self.world, 1
            elif self.world % A == 0:
                self.n_groups, self.sub_world = A, self.world // A
            else:
                raise ValueError(f"shard='accumulation' needs world ({self.world}) to divide or be a multiple of "
                                 f"n_grad_accum_itrs ({A})")
            if self.sub_world > 1:  # every rank creates every group, in the same order
                for g in range(self.n_groups):
                    grp = dist.new_group(list(range(g * self.sub_world, (g + 1) * self.sub_world)))
                    if g == self.rank // self.sub_world:
                        self.sub_group = grp
        self.group_index, self.sub_rank = self.rank // self.sub_world, self.rank % self.sub_world
        lo, hi = shard_bounds(batch_size, self.sub_rank, self.sub_world)
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
        self.standardize_global = standardize_global and self.sub_world > 1
        self.device = next(model.parameters()).device
        self.use_cuda_graph = bool(use_cuda_graph)
        self.log_every = max(1, int(log_every))
        self._schedule = None if disable_scheduler else self.scheduler._factor
        if self.use_cuda_graph:
            self._init_graph_mode(lr)

    # ------------------------------------------------------------------ pieces
    def _allreduce_grads(self):
        if self.world == 1:
            return
        # every rank reduces EVERY parameter: a rank whose whole accumulation window kept no sample has no .grad
        # tensors yet and contributes zeros (a rank-dependent parameter list would mismatch the collective)
        for p in self.model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.model.parameters()]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)  # NCCL over NVLink; ~45 MB for ResNet-18
        offset = 0
        for g in grads:
            g.copy_(flat[offset:offset + g.numel()].view_as(g))
            offset += g.numel()

    # ------------------------------------------------------------------ the iteration
    def _draw(self, itr):
        """The random inputs of an iteration: (subject index, bone contrast, Euler angles in degrees (B,3),
        translations (B,3)).  Subject and contrast come from the stream every rank shares, the poses from the
        rank's own; tests replay recorded draws through this hook."""
        subject = int(torch.randint(len(self.volumes), (1,), generator=self.shared_rng))
        contrast = float(torch.empty(1).uniform_(1.0, 10.0, generator=self.shared_rng))
        rng = self.pose_rng
        if self.n_groups > 1:  # iterations run on different groups: the pose stream is a function of the iteration
            rng = torch.Generator().manual_seed((self.seed * 9973 + 1 + self.sub_rank) * 1_000_003 + itr)
        rot, xyz = random_pose_params(**self.pose_distribution, batch_size=self.local_batch, generator=rng)
        return subject, contrast, rot, xyz

    def owns(self, itr):
        """Does this rank's group work on iteration ``itr``?  (Always, unless the accumulation window is sharded.)"""
        return (itr % self.n_grad_accum_itrs) % self.n_groups == self.group_index

    def step(self, itr):
        if self.use_cuda_graph:
            return self._step_graphed(itr)
        if self.world > 1:
            return self._step_masked_eager(itr)
        dev = self.device
        subject, contrast, rot, xyz = self._draw(itr)
        vol, seg, affinv, offset = self.volumes[subject]

        pose = convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY", degrees=True)
        pose = pose.compose(offset)

        density = transform_hu_to_density(vol, contrast)
        with torch.no_grad():
            img, mask, keep = render_samples(self.drr, density, seg, affinv, pose)
        img, mask, pose = img[keep], mask[keep], pose[keep]

        # single rank: the reference's dynamic shapes (samples that miss the volume are indexed away).  Several ranks
        # never come here: a rank-local `if len(pose) > 0` around collectives would mismatch them whenever one rank
        # keeps nothing, so they run the masked, static-shape iteration (_step_masked_eager) instead.
        kept = torch.tensor([float(keep.sum())], device=dev)
        log = {"kept": kept.item() / self.batch_size}
        if len(pose) > 0:
            x = self.transforms(img)
            pred_pose = self.model(x)
            pred_img, pred_mask, _ = render_samples(self.drr, density, seg, affinv, pred_pose)
            x_true, x_pred = x, self.transforms(pred_img)
            loss, mncc, dgeo, rgeo, tgeo, dice, mvc = self.lossfn(x_true, mask, pose, x_pred, pred_mask, pred_pose)
            (loss.sum() / kept.clamp_min(1.0) / self.n_grad_accum_itrs).backward()
            sums = torch.stack([loss.sum(), mncc.sum(), dgeo.sum(), rgeo.sum(), tgeo.sum(), dice.sum()]).detach()
        else:
            sums = torch.zeros(6, device=dev)
        means = (sums / kept.clamp_min(1.0)).tolist()
        means[0] /= self.n_grad_accum_itrs  # the reference logs the loss after dividing it for accumulation
        log.update(dict(zip(("loss", "mncc", "dgeo", "rgeo", "tgeo", "dice"), means)))

        if (itr + 1) % self.n_grad_accum_itrs == 0 or (itr + 1) == self.n_total_itrs:
            self._allreduce_grads()
            adaptive_clip_grad_(self.model.parameters())
            self.optimizer.step()
            self.scheduler.step()
            self.optimizer.zero_grad()
        log["lr"] = self.scheduler.get_last_lr()[0]
        return log

    def checkpoint(self, path, itr, config, model_number=0):
        """Write a ``*.pth`` with the reference's keys (trainer.py:318-332); rank 0 only, every rank may call it."""
        from .inference import save_checkpoint  # noqa: PLC0415

        if self.rank == 0:
            optimizer, scheduler = self.optimizer, self.scheduler
            if self.use_cuda_graph:
                optimizer, scheduler = self._portable_optimizer_state()
            save_checkpoint(path, self.model, optimizer, scheduler, itr, model_number, config)
        return path

    def _portable_optimizer_state(self):
        """Graph mode drives a capturable Adam with a device-side learning rate by hand; the reference's
        ``reuse_optimizer`` resume expects the state of a plain Adam + its scheduler.  Returns objects whose
        ``state_dict()`` is exactly that: float learning rates, CPU step counters, and a scheduler whose counters
        (``last_epoch``, ``_last_lr``, ``_step_count``) stand where ``_opt_steps`` optimiser steps leave them."""
        lr_now = self._base_lr * self._lr_factor(self._opt_steps)
        sd = self.optimizer.state_dict()
        for g in sd["param_groups"]:
            g["lr"] = lr_now
            g["capturable"] = False
            g["initial_lr"] = self._base_lr  # what a scheduler attached to the optimizer records
        for st in sd["state"].values():
            if torch.is_tensor(st.get("step")):
                st["step"] = st["step"].detach().to("cpu", torch.float32)
        self.scheduler.last_epoch = self._opt_steps
        self.scheduler._step_count = self._opt_steps + 1
        self.scheduler._last_lr = [lr_now for _ in self.scheduler.optimizer.param_groups]

        class _State:  # state_dict() carrier
            def __init__(self, d):
                self._d = d

            def state_dict(self):
                return self._d

        return _State(sd), self.scheduler

    # ------------------------------------------------------------------ the iteration as CUDA graphs
    # The eager iteration above costs ~20 ms of host time (a ResNet forward/backward, two renders, ~10^3 small
    # launches and five host round trips), more than its GPU time -- and at 8 ranks the GPU share shrinks 8x while
    # the host share does not.  Graph mode replays one captured graph per subject for the iteration and one for the
    # optimiser step.  What makes it capturable:
    #   * static shapes: instead of dropping the samples that miss the volume (img[keep], trainer.py:202-204) they
    #     stay in the batch with weight 0.  Exact, not an approximation: the network normalises per sample
    #     (GroupNorm), the losses are per sample, the batch-global min/max of Standardize is taken over the kept
    #     samples only, and the mean divides by the kept count;
    #   * no host reads: kept count and log values stay on the device and are fetched every ``log_every`` steps;
    #   * run-time scalars in device memory: pose batch, contrast (density kernel reads it), learning rate
    #     (Adam(capturable=True) with a tensor lr, set from the same WarmupCosineSchedule formula);
    #   * the density goes to a static buffer and its texture upload is part of the captured work;
    #   * collectives (kept count, log sums, min/max, gradient all-reduce) are captured NCCL calls.
    def _standardize_masked(self, x, keep, stats=None):
        """XrayTransforms whose batch-global min/max run over the KEPT samples only (the reference standardises
        img[keep]); with nothing kept anywhere the range falls back to [0, 1] and every weight is 0 anyway.
        ``stats`` (B,4) from the render epilogue carries the per-sample min / max already: the batch-global range is
        then a reduction over B numbers, and the ranks exchange it in ONE collective (max of [-lo, hi])."""
        if stats is not None:
            lo = torch.where(keep, stats[:, 2], torch.full_like(stats[:, 2], float("inf"))).min()
            hi = torch.where(keep, stats[:, 3], torch.full_like(stats[:, 3], float("-inf"))).max()
        else:
            sel = keep.view(-1, 1, 1, 1)
            lo = torch.where(sel, x.detach(), torch.full_like(x, float("inf"))).min()
            hi = torch.where(sel, x.detach(), torch.full_like(x, float("-inf"))).max()
        if self.standardize_global:
            lohi = torch.stack([-lo, hi])
            dist.all_reduce(lohi, op=dist.ReduceOp.MAX, group=self.sub_group)
            lo, hi = -lohi[0], lohi[1]
        lo = torch.where(torch.isfinite(lo), lo, torch.zeros_like(lo))
        hi = torch.where(torch.isfinite(hi), hi, torch.ones_like(hi))
        t = self.transforms
        x = (x - lo) / (hi - lo + t.standardize.eps)
        if t.equalize is not None:
            x = t.equalize(x)
        if tuple(x.shape[-2:]) != t.size:
            x = torch.nn.functional.interpolate(x, size=t.size, mode="bilinear", align_corners=False, antialias=True)
        return (x - t.mean) / t.std

    def _step_masked_eager(self, itr):
        """The eager iteration of a multi-rank run: static shapes (dropped samples keep their slot with weight 0), so
        that every rank issues the same collectives whatever it kept -- the same arithmetic as graph mode, without
        the capture."""
        dev = self.device
        subject, contrast, rot, xyz = self._draw(itr)  # every rank draws every iteration: the shared stream stays in step
        log_dev = None
        if self.owns(itr):
            log_dev = self._masked_iteration(subject, rot.to(dev), xyz.to(dev), contrast, None)
            self._last_eager_log = log_dev
        if (itr + 1) % self.n_grad_accum_itrs == 0 or (itr + 1) == self.n_total_itrs:
            self._allreduce_grads()
            adaptive_clip_grad_(self.model.parameters())
            self.optimizer.step()
            self.scheduler.step()
            self.optimizer.zero_grad()
        log_dev = log_dev if log_dev is not None else getattr(self, "_last_eager_log", None)
        vals = log_dev.tolist() if log_dev is not None else [float("nan")] * 7
        log = dict(zip(("loss", "mncc", "dgeo", "rgeo", "tgeo", "dice", "kept"), vals))
        log["lr"] = self.scheduler.get_last_lr()[0]
        return log

    def _device_iteration(self, subject):
        vol = self.volumes[subject][0]
        self._log.copy_(self._masked_iteration(subject, self._rot, self._xyz, self._contrast,
                                               self._density_buffer(vol)))

    def _masked_iteration(self, subject, rot, xyz, contrast, density_out):
        """One iteration with static shapes; returns the 7 log values (loss, mncc, dgeo, rgeo, tgeo, dice, kept
        fraction) as a device tensor.  ``contrast`` is a float or a 1-element device tensor."""
        vol, seg, affinv, offset = self.volumes[subject]
        pose = convert(rot, xyz, parameterization="euler_angles", convention="ZXY", degrees=True)
        pose = pose.compose(offset)
        density = transform_hu_to_density(vol, contrast, out=density_out)
        if density_out is not None and hasattr(self.drr.renderer, "_texture"):
            self.drr.renderer._texture.invalidate()  # the buffer is rewritten through its raw pointer: upload again
        with torch.no_grad():
            img, mask, keep, stats = render_samples_fused(self.drr, density, seg, affinv, pose)
        w = keep.to(torch.float32)
        kept = w.sum().reshape(1)
        if self.sub_world > 1:
            dist.all_reduce(kept, op=dist.ReduceOp.SUM, group=self.sub_group)
        x = self._standardize_masked(img, keep, stats)
        pred_pose = self.model(x)
        pred_img, pred_mask, _, pred_stats = render_samples_fused(self.drr, density, seg, affinv, pred_pose)
        x_pred = self._standardize_masked(pred_img, keep, pred_stats)
        if mask is None:  # no label channels: the Dice term is the constant 1 (DiceLoss on one-channel masks)
            mask = pred_mask = torch.ones(len(w), 1, 1, 1, device=w.device, dtype=torch.bool)
        loss, mncc, dgeo, rgeo, tgeo, dice, _ = self.lossfn(x, mask, pose, x_pred, pred_mask, pred_pose)
        denom = kept.clamp_min(1.0)
        ((loss * w).sum() / denom / self.n_grad_accum_itrs).backward()
        with torch.no_grad():
            sums = torch.stack([(v.detach() * w).sum() for v in (loss, mncc, dgeo, rgeo, tgeo, dice)])
            if self.sub_world > 1:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.sub_group)
            means = sums / denom
            means[0] /= self.n_grad_accum_itrs  # the reference logs the loss after dividing it for accumulation
            return torch.cat([means, kept / self.batch_size])

    def _device_optimizer_step(self):
        self._allreduce_grads()
        adaptive_clip_grad_(self.model.parameters())
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=False)

    def _density_buffer(self, vol):
        key = tuple(vol.shape)
        if key not in self._density:
            self._density[key] = torch.empty(key, device=vol.device, dtype=torch.float32)
        return self._density[key]

    def _init_graph_mode(self, lr):
        if self.lossfn.weight_mvc > 0:
            raise NotImplementedError("use_cuda_graph: the multiview-consistency term pairs samples across the batch "
                                      "and is not masked; train with weight_mvc = 0 (the reference default)")
        dev = self.device
        B = self.local_batch
        self._rot = torch.zeros(B, 3, device=dev)
        self._xyz = torch.zeros(B, 3, device=dev)
        self._contrast = torch.ones(1, device=dev)
        self._log = torch.zeros(7, device=dev)
        self._density = {}
        self._graphs = {}
        self._opt_graph = None
        self._pool = None
        self._opt_steps = 0
        self._base_lr = lr
        self._lr = torch.tensor(lr * self._lr_factor(0), device=dev, dtype=torch.float32)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self._lr, capturable=True)
        # the scheduler object is only a carrier of the state the checkpoint needs (never stepped in graph mode):
        # keep it on a plain Adam so that its state_dict is the one the reference's resume expects
        carrier = torch.optim.Adam(self.model.parameters(), lr=lr)
        if self._schedule is None:
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(carrier, lambda step: 1.0)
        else:
            self.scheduler = WarmupCosineSchedule(carrier, self.scheduler.warmup_steps, self.scheduler.t_total)
        for p in self.model.parameters():  # static .grad tensors shared by every graph
            p.grad = torch.zeros_like(p)
        # pinned staging ring for the per-step pose batch: the host may run several steps ahead of the device
        self._stage = [(torch.empty(B, 3).pin_memory(), torch.empty(B, 3).pin_memory(), torch.cuda.Event())
                       for _ in range(8)]
        self._stage_used = [False] * 8
        self._last_log = {}

    def _lr_factor(self, opt_step):
        return 1.0 if self._schedule is None else self._schedule(opt_step)

    def _capture(self, fn, *args):
        """Warm up on a side stream (lazy initialisation, allocator pools, texture creation), then capture."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(*args)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self._pool):
            fn(*args)
        if self._pool is None:
            self._pool = graph.pool()
        return graph

    def _step_graphed(self, itr):
        subject, contrast, rot, xyz = self._draw(itr)
        if not self.owns(itr):  # another group's iteration of the accumulation window
            return self._finish_graphed(itr)
        slot = itr % len(self._stage)
        rot_h, xyz_h, ev = self._stage[slot]
        if self._stage_used[slot]:
            ev.synchronize()  # the copies that last read these pinned buffers have run
        rot_h.copy_(rot)
        xyz_h.copy_(xyz)
        self._rot.copy_(rot_h, non_blocking=True)
        self._xyz.copy_(xyz_h, non_blocking=True)
        ev.record()
        self._stage_used[slot] = True
        self._contrast.fill_(contrast)

        if subject not in self._graphs:
            # warm-up + capture run the iteration twice without it counting: keep the accumulated gradients
            saved = [p.grad.clone() for p in self.model.parameters()]
            self._graphs[subject] = self._capture(self._device_iteration, subject)
            for p, g in zip(self.model.parameters(), saved):
                p.grad.copy_(g)
        self._graphs[subject].replay()
        return self._finish_graphed(itr)

    def _finish_graphed(self, itr):
        if (itr + 1) % self.n_grad_accum_itrs == 0 or (itr + 1) == self.n_total_itrs:
            if self._opt_graph is None:
                self._opt_graph = self._capture_optimizer()
            else:
                self._opt_graph.replay()
            self._opt_steps += 1
            self._lr.fill_(self._base_lr * self._lr_factor(self._opt_steps))
        if (itr + 1) % self.log_every == 0:
            vals = self._log.tolist()  # the only host round trip, every log_every iterations
            self._last_log = dict(zip(("loss", "mncc", "dgeo", "rgeo", "tgeo", "dice", "kept"), vals))
        log = dict(self._last_log)
        log["lr"] = self._base_lr * self._lr_factor(self._opt_steps)
        return log

    def _capture_optimizer(self):
        """The first optimiser step runs eagerly on a side stream (Adam's lazy state initialisation must not be
        captured) and counts; the graph captured right after it serves every later step.  Capturing executes
        nothing, but the eager warm-up inside ``_capture`` would: so capture by hand here."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._device_optimizer_step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self._pool):
            self._device_optimizer_step()
        return graph
