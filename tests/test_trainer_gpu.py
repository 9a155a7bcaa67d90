"""The training iteration (trainer.py:185-246) on the CUDA path: runs, learns, and its loss normalisation makes
sharded gradients add up to the unsharded ones."""

import numpy as np
import pytest
import torch

import xvr_b200
from tests._scene import POSE_RANGES, SDD, pixel_size
from xvr_b200.data import read, synthetic_ct
from xvr_b200.pose import RigidTransform, convert
from xvr_b200.preprocess import XrayTransforms
from xvr_b200.trainer import DiceLoss, PoseRegressionLoss, PoseRegressor, TrainStep, adaptive_clip_grad_, render_samples

pytestmark = pytest.mark.gpu


def _setup(cuda, n=64, h=32, labels=False, n_vols=2):
    volumes, drr = [], None
    for seed in range(n_vols):
        hu, lab, affine = synthetic_ct(n, seed=seed, with_labels=labels, device=cuda)
        sub = read(hu, lab, affine=affine, center_volume=False)
        if drr is None:
            drr = xvr_b200.DRR(sub, SDD, h, pixel_size(h), renderer="trilinear", reverse_x_axis=False).to(cuda)
            drr.density = None
        aff = torch.as_tensor(affine, dtype=torch.float32, device=cuda)
        center = aff[:3, :3] @ ((torch.tensor(hu.shape, device=cuda) - 1) / 2) + aff[:3, 3]
        offset = convert(torch.zeros(1, 3, device=cuda), center[None], parameterization="euler_angles", convention="ZXY")
        volumes.append((hu, lab.float() if labels else None, RigidTransform(torch.linalg.inv(aff)), offset))
    torch.manual_seed(0)
    model = PoseRegressor("resnet18", "quaternion_adjugate", "ZXY", height=h).to(cuda)
    return drr, model, volumes


@pytest.mark.parametrize("labels", [False, True])
def test_training_iterations_run_and_update_the_model(cuda, labels):
    drr, model, volumes = _setup(cuda, labels=labels)
    step = TrainStep(drr, model, volumes, POSE_RANGES, XrayTransforms(32), SDD, batch_size=6, n_grad_accum_itrs=2,
                     n_warmup_itrs=2, lr=1e-3)
    before = [p.detach().clone() for p in model.parameters()]
    logs = [step.step(i) for i in range(4)]
    for log in logs:
        assert all(np.isfinite(v) for v in log.values()), log
        assert 0.0 <= log["kept"] <= 1.0
    assert logs[0]["kept"] > 0.5
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_sharded_loss_normalisation_reproduces_unsharded_gradient(cuda):
    """d/dtheta of (sum of a shard's losses / global kept count), summed over shards == d/dtheta of the global mean."""
    drr, model, volumes = _setup(cuda)
    lossfn = PoseRegressionLoss(SDD)
    vol, seg, affinv, offset = volumes[0]
    from xvr_b200.sampler import random_pose_params

    rot, xyz = random_pose_params(**POSE_RANGES, batch_size=6, generator=torch.Generator().manual_seed(3))
    pose = convert(rot.to(cuda), xyz.to(cuda), parameterization="euler_angles", convention="ZXY", degrees=True).compose(offset)
    density = xvr_b200.transform_hu_to_density(vol, 2.0)
    with torch.no_grad():
        img, mask, keep = render_samples(drr, density, seg, affinv, pose)
    assert keep.all()
    lo, hi = img.min(), img.max()

    def grads(sl):
        model.zero_grad()
        x = ((img[sl] - lo) / (hi - lo + 1e-6) - 0.15) / 0.1
        pred = model(x)
        pimg, pmask, _ = render_samples(drr, density, seg, affinv, pred)
        xp = ((pimg - lo) / (hi - lo + 1e-6) - 0.15) / 0.1
        loss = lossfn(x, mask[sl], pose[sl], xp, pmask, pred)[0]
        (loss.sum() / 6.0).backward()
        return torch.cat([p.grad.flatten() for p in model.parameters()])

    full = grads(slice(0, 6))
    parts = grads(slice(0, 3)) + grads(slice(3, 6))
    assert ((full - parts).norm() / full.norm()).item() < 1e-3


def test_dice_and_agc_semantics(cuda):
    a = torch.zeros(2, 3, 4, 4, device=cuda, dtype=torch.bool)
    b = torch.zeros_like(a)
    a[:, 1, :2], b[:, 1, :2] = True, True      # identical channel 1 -> dice 1
    a[0, 2, 0, 0] = True                        # channel 2 disjoint in sample 0 -> dice 0; empty in sample 1 -> nan, ignored
    b[0, 2, 3, 3] = True
    loss = DiceLoss()(a, b)
    assert torch.allclose(loss, torch.tensor([0.5, 0.0], device=cuda))
    p = torch.nn.Parameter(torch.ones(4, 8, device=cuda))
    p.grad = torch.full_like(p, 10.0)
    adaptive_clip_grad_([p], clip_factor=0.01, eps=1e-3)
    assert torch.allclose(p.grad.norm(dim=1), torch.full((4,), 0.01 * 8**0.5, device=cuda), rtol=1e-4)


@pytest.mark.parametrize("wide,labels", [(False, False), (True, False), (False, True)])
def test_graph_mode_matches_eager(cuda, wide, labels):
    """The CUDA-graph iteration keeps dropped samples in the batch with weight 0 instead of indexing them away
    (static shapes); it must train like the eager, reference-shaped iteration: same kept fraction, same losses,
    same parameter trajectory (up to the round-off of a different batch size inside cuDNN/cuBLAS)."""
    ranges = dict(POSE_RANGES)
    if wide:  # some poses lose the volume -> keep < 1
        ranges.update(txmin=-450.0, txmax=450.0, tzmin=-450.0, tzmax=450.0)
    logs, final, init = {}, {}, None
    for graph in (False, True):
        drr, model, volumes = _setup(cuda, labels=labels)
        init = torch.cat([p.detach().flatten().clone() for p in model.parameters()])
        step = TrainStep(drr, model, volumes, ranges, XrayTransforms(32), SDD, batch_size=8, n_grad_accum_itrs=2,
                         n_warmup_itrs=2, lr=1e-3, use_cuda_graph=graph)
        logs[graph] = [step.step(i) for i in range(6)]
        final[graph] = torch.cat([p.detach().flatten() for p in model.parameters()])
    kept = [log["kept"] for log in logs[False]]
    if wide:
        assert min(kept) < 1.0 and max(kept) > 0.0, kept
    for i, (a, b) in enumerate(zip(logs[False], logs[True])):
        assert a["kept"] == pytest.approx(b["kept"], abs=1e-6), (i, a, b)
        assert a["lr"] == pytest.approx(b["lr"], rel=1e-6), (i, a, b)
        for k in ("loss", "mncc", "dgeo", "dice"):
            assert a[k] == pytest.approx(b[k], rel=5e-3 if i < 2 else 5e-2, abs=1e-3), (i, k, a, b)
    da, db = final[False] - init, final[True] - init
    assert da.norm() > 0
    assert (torch.dot(da, db) / (da.norm() * db.norm())).item() > 0.97


class _TinyBackbone(torch.nn.Sequential):
    """Same layers as tests/golden/make_reference_train_golden.py hands the reference through its timm stub."""

    def __init__(self):
        super().__init__(torch.nn.Conv2d(1, 4, 3, stride=2, padding=1), torch.nn.GroupNorm(2, 4), torch.nn.ReLU(),
                         torch.nn.Conv2d(4, 8, 3, stride=2, padding=1), torch.nn.GroupNorm(2, 8), torch.nn.ReLU(),
                         torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("run", ["plain", "labels"])
def test_train_step_follows_the_genuine_reference_step(cuda, run, graph):
    """tests/golden/reference_train_v1.pt: xvr's own Trainer.step / render_samples / load (unmodified) on the oracle
    renderer, six iterations with gradient accumulation, pose ranges wide enough that most batches lose samples.
    Replaying its random draws, our iteration -- eager with dynamic shapes, and graph-captured with masked static
    shapes -- must keep the same samples, log the same losses and move the CNN to the same weights."""
    import os

    from tests.golden.make_golden import scene

    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_train_v1.pt"), weights_only=False)
    sc, g = gold["scene"], gold["runs"][run]
    hu, labels, affine = scene(sc["n"])
    hu = hu.to(cuda)
    seg = labels.to(cuda).float() if run == "labels" else None
    sub = read(hu, labels.to(cuda) if run == "labels" else None, affine=affine, center_volume=False)
    drr = xvr_b200.DRR(sub, sc["sdd"], sc["height"], sc["delx"], renderer="trilinear", reverse_x_axis=False).to(cuda)
    drr.density = None
    aff = torch.as_tensor(affine, dtype=torch.float32, device=cuda)
    center = aff[:3, :3] @ ((torch.tensor(hu.shape, device=cuda) - 1) / 2) + aff[:3, 3]
    offset = convert(torch.zeros(1, 3, device=cuda), center[None], parameterization="euler_angles", convention="ZXY")
    volumes = [(hu, seg, RigidTransform(torch.linalg.inv(aff)), offset)]
    model = PoseRegressor("tiny", "euler_angles", "ZXY", height=sc["height"], backbone=_TinyBackbone()).to(cuda)
    model.load_state_dict(g["init_state"])
    step = TrainStep(drr, model, volumes, gold["ranges"], XrayTransforms(sc["height"]), sc["sdd"],
                     batch_size=gold["batch"], lr=gold["lr"], n_total_itrs=1000, n_warmup_itrs=gold["warmup"],
                     n_grad_accum_itrs=gold["accum"], use_cuda_graph=graph, **gold["weights"])
    draws = iter(g["draws"])

    def replay(itr):
        d = next(draws)
        return 0, d["contrast"], d["rot_xyz_deg"][:, :3].clone(), d["rot_xyz_deg"][:, 3:].clone()

    step._draw = replay
    for itr, ref in enumerate(g["logs"]):
        log = step.step(itr)
        assert log["kept"] == pytest.approx(ref["kept"], abs=1e-6), (itr, log, ref)
        assert log["lr"] == pytest.approx(ref["lr"], rel=1e-6), (itr, log, ref)
        for k in ("loss", "mncc", "dgeo", "rgeo", "tgeo", "dice"):
            assert log[k] == pytest.approx(ref[k], rel=2e-2, abs=2e-3), (itr, k, log, ref)
    init = torch.cat([v.flatten() for v in g["init_state"].values()])
    want = torch.cat([v.flatten() for v in g["final_state"].values()]) - init
    got = torch.cat([v.detach().flatten().cpu() for v in model.state_dict().values()]) - init
    assert want.norm() > 0
    assert (torch.dot(got, want) / (got.norm() * want.norm())).item() > 0.98
    assert (got - want).norm().item() < 0.15 * want.norm().item()
