"""The genuine training iteration of xvr (``Trainer.step`` / ``render_samples`` / ``load``,
/root/reference/src/xvr/model/trainer.py:185-304, with the genuine ``PoseRegressor.forward``, ``PoseRegressionLoss``,
``get_random_pose``, ``XrayTransforms`` and ``WarmupCosineSchedule``) run on the ORACLE renderer and metrics.

    python tests/golden/make_reference_train_golden.py        # needs /root/reference; writes reference_train_v1.pt

What this pins: everything the iteration itself owns -- the draw order (poses, then contrast), composing the poses
with the isocentre offset, the keep rule and what it filters, standardising images and predictions separately, the
loss composition and its weights, ``loss / n_grad_accum_itrs`` then ``.mean().backward()``, when the optimiser /
scheduler step and what the log holds.  What it does not pin: renderer, similarity and pose arithmetic (oracle,
through adapters shaped like DiffDRR's ``RigidTransform`` / ``DRR`` / metric classes), the timm backbone (a tiny
CNN handed out by a stub ``timm.create_model``), timm's ``adaptive_clip_grad`` (our restatement is injected: AGC
stays unpinned) and the kornia augmentations (identity).  The xvr files are loaded unmodified; ``.cuda()`` is
patched to a no-op for the duration of the run (the build container has no GPU).
"""

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/xvr"
OUT = os.path.join(HERE, "reference_train_v1.pt")

import oracle  # noqa: E402
from tests.golden.make_golden import scene  # noqa: E402

N_VOL, HEIGHT, SDD = 32, 24, 1020.0
DELX = 1.08821875 * 256.0 / HEIGHT
BATCH, N_ITRS, ACCUM, WARMUP, LR = 6, 6, 2, 4, 1e-3
RANGES = dict(alphamin=-30.0, alphamax=30.0, betamin=-30.0, betamax=30.0, gammamin=-10.0, gammamax=10.0,
              txmin=-260.0, txmax=260.0, tymin=700.0, tymax=900.0, tzmin=-260.0, tzmax=260.0)
WEIGHTS = dict(weight_ncc=1.0, weight_geo=1e-2, weight_dice=1.0, weight_mvc=0.0)


class RT:
    """RigidTransform-shaped adapter over oracle's (B,4,4) helpers."""

    def __init__(self, matrix):
        self.matrix = matrix

    def __len__(self):
        return len(self.matrix)

    def __getitem__(self, idx):
        return RT(self.matrix[idx])

    def __matmul__(self, other):
        return RT(self.matrix @ other.matrix)

    def __call__(self, pts):
        return oracle.apply(self.matrix, pts)

    def inverse(self):
        return RT(oracle.invert(self.matrix))

    def compose(self, other):
        return RT(oracle.compose(self.matrix, other.matrix))

    def convert(self, parameterization, convention=None):
        return oracle.params_from_pose(self.matrix, parameterization, convention)

    def cuda(self):
        return self


def convert(rot, xyz, parameterization, convention=None, degrees=False):
    return RT(oracle.pose_from_params(rot, xyz, parameterization, convention, degrees))


class MNCC(torch.nn.Module):
    def __init__(self, patch_sizes, patch_weights):
        super().__init__()
        self.p, self.w = tuple(patch_sizes), tuple(patch_weights)

    def forward(self, a, b):
        return oracle.multiscale_ncc(a, b, self.p, self.w)


class Geodesic(torch.nn.Module):
    def __init__(self, sdd, eps=None):
        super().__init__()
        self.sdd, self.eps = sdd, eps

    def forward(self, a, b):
        return oracle.double_geodesic(a.matrix, b.matrix, self.sdd, self.eps)


class TinyBackbone(torch.nn.Sequential):
    def __init__(self):
        super().__init__(torch.nn.Conv2d(1, 4, 3, stride=2, padding=1), torch.nn.GroupNorm(2, 4), torch.nn.ReLU(),
                         torch.nn.Conv2d(4, 8, 3, stride=2, padding=1), torch.nn.GroupNorm(2, 8), torch.nn.ReLU(),
                         torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())


def load_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    def load(rel, name, package):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        m.__package__ = package
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    from xvr_b200.trainer import adaptive_clip_grad_  # our restatement of timm's AGC (unpinned by this golden)

    none = lambda *a, **k: None  # noqa: E731
    stubs = {
        "matplotlib": mod("matplotlib"), "matplotlib.pyplot": mod("matplotlib.pyplot"), "wandb": mod("wandb"),
        "timm": mod("timm", create_model=lambda *a, **k: TinyBackbone()), "timm.utils": mod("timm.utils"),
        "timm.utils.agc": mod("timm.utils.agc", adaptive_clip_grad=adaptive_clip_grad_),
        "diffdrr": mod("diffdrr"),
        "diffdrr.data": mod("diffdrr.data", transform_hu_to_density=oracle.hu_to_density),
        "diffdrr.pose": mod("diffdrr.pose", RigidTransform=RT, convert=convert),
        "diffdrr.metrics": mod("diffdrr.metrics", DoubleGeodesicSE3=Geodesic, MultiscaleNormalizedCrossCorrelation2d=MNCC),
        "diffdrr.registration": mod("diffdrr.registration", N_ANGULAR_COMPONENTS=oracle.N_ANGULAR_COMPONENTS),
        "diffdrr.visualization": mod("diffdrr.visualization", plot_drr=None, plot_mask=None),
        "xvr": mod("xvr", __path__=[]), "xvr.model": mod("xvr.model", __path__=[]),
        "xvr.config": mod("xvr.config", __path__=[]),
        "xvr.model.augmentations": mod("xvr.model.augmentations", XrayAugmentations=none),
        "xvr.model.utils": mod("xvr.model.utils", initialize_coordinate_frame=none, initialize_modules=none,
                               initialize_subjects=none),
    }
    saved = {k: sys.modules.get(k) for k in list(stubs) + ["xvr.config.trainer", "xvr.model.loss", "xvr.model.sampler",
                                                           "xvr.model.network", "xvr.model.trainer", "xvr.model.scheduler"]}
    sys.modules.update(stubs)
    try:
        load("config/trainer.py", "xvr.config.trainer", "xvr.config")
        load("model/loss.py", "xvr.model.loss", "xvr.model")
        load("model/sampler.py", "xvr.model.sampler", "xvr.model")
        sched = load("model/scheduler.py", "xvr.model.scheduler", "xvr.model")
        net = load("model/network.py", "xvr.model.network", "xvr.model")
        trainer = load("model/trainer.py", "xvr.model.trainer", "xvr.model")
        loss = sys.modules["xvr.model.loss"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    pre_spec = importlib.util.spec_from_file_location("_ref_preprocess", os.path.join(REF, "utils/preprocess.py"))
    pre = importlib.util.module_from_spec(pre_spec)
    pre_spec.loader.exec_module(pre)
    return trainer, net, loss, sched, pre


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    trainer, net, loss, sched, pre = load_modules()
    hu, labels, affine = scene(N_VOL)
    affinv = torch.as_tensor(np.linalg.inv(affine), dtype=torch.float32)[None]
    aff = torch.as_tensor(affine, dtype=torch.float32)
    center = (aff[:3, :3] @ ((torch.tensor(hu.shape, dtype=torch.float32) - 1) / 2) + aff[:3, 3])[None]

    def detector(pose, calibration):
        return oracle.detector_rays(pose.matrix, oracle.REORIENT["AP"], HEIGHT, HEIGHT, DELX, DELX, 0.0, 0.0, SDD, False)

    def renderer(vol, source, target, img, mask=None):
        return oracle.trilinear_render(vol, source, target, img, mask=mask)

    goldens = {}
    for name, seg in (("plain", None), ("labels", labels.to(torch.float32))):
        torch.manual_seed(11)
        model = net.PoseRegressor("tiny", "euler_angles", "ZXY", height=HEIGHT, unit_conversion_factor=1000.0)
        with torch.no_grad():  # start near the pose range so that predicted DRRs see the volume
            model.xyz_regression.bias.copy_(torch.tensor([0.0, 0.8, 0.0]))
            model.xyz_regression.weight.mul_(0.05)
            model.rot_regression.weight.mul_(0.2)
        init_state = {k: v.clone() for k, v in model.state_dict().items()}
        optimizer = torch.optim.Adam(model.parameters(), lr=LR)
        scheduler = sched.WarmupCosineSchedule(optimizer, WARMUP / ACCUM, 1000 / ACCUM)
        drr = types.SimpleNamespace(
            volume=hu, mask=seg, affine_inverse=RT(affinv), center=center, detector=detector, renderer=renderer,
            reshape_transform=lambda img, batch_size: img.view(batch_size, -1, HEIGHT, HEIGHT))
        me = types.SimpleNamespace(
            pose_distribution=dict(RANGES, batch_size=BATCH), contrast_distribution=torch.distributions.Uniform(1.0, 10.0),
            drr=drr, transforms=pre.XrayTransforms(HEIGHT), augmentations=lambda x: x, model=model, reframe=None,
            lossfn=loss.PoseRegressionLoss(SDD, **WEIGHTS), n_grad_accum_itrs=ACCUM, n_total_itrs=1000,
            optimizer=optimizer, scheduler=scheduler)
        me.load = types.MethodType(trainer.Trainer.load, me)
        me.render_samples = types.MethodType(trainer.Trainer.render_samples, me)

        draws, logs = [], []
        real_pose, real_hu = trainer.get_random_pose, trainer.transform_hu_to_density

        def spy_pose(**kw):
            pose = real_pose(**kw)
            draws.append({"rot_xyz_deg": torch.cat(oracle.params_from_pose(pose.matrix, "euler_angles", "ZXY", degrees=True), -1)})
            return pose

        def spy_hu(vol, contrast):
            draws[-1]["contrast"] = contrast
            return real_hu(vol, contrast)

        trainer.get_random_pose, trainer.transform_hu_to_density = spy_pose, spy_hu
        patched = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            torch.manual_seed(5)
            for itr in range(N_ITRS):
                log, imgs, masks = trainer.Trainer.step(me, itr, None)
                logs.append(log)
        finally:
            torch.Tensor.cuda = patched
            trainer.get_random_pose, trainer.transform_hu_to_density = real_pose, real_hu
        goldens[name] = {"init_state": init_state, "draws": draws, "logs": logs,
                         "final_state": {k: v.detach().clone() for k, v in model.state_dict().items()}}
        print(name, [round(float(l["kept"]), 3) for l in logs], [round(float(l["loss"]), 4) for l in logs])

    out = {"source": "src/xvr/model/trainer.py:185-304 (Trainer.step, load, render_samples), unmodified",
           "scene": dict(n=N_VOL, height=HEIGHT, delx=DELX, sdd=SDD), "batch": BATCH, "accum": ACCUM, "warmup": WARMUP,
           "lr": LR, "ranges": RANGES, "weights": WEIGHTS, "runs": goldens}
    torch.save(out, OUT)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    main()
