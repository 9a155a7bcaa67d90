"""Genuine xvr modules executed on top of the alias package (``import diffdrr`` -> xvr_b200), next to their mirrors.

Only where /root/reference is present (the build container; the GPU boxes do not have it): these tests show the
drop-in claim itself -- unmodified xvr code, loaded by path, runs against this package -- and pin the mirrors'
own arithmetic to the genuine code's on the same inputs."""

import importlib.util
import os
import sys

import pytest
import torch

import xvr_b200

REF = "/root/reference/src/xvr"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not available here")


@pytest.fixture
def diffdrr_alias(monkeypatch):
    import xvr_b200.compat

    monkeypatch.syspath_prepend(os.path.dirname(xvr_b200.compat.__file__))
    for name in [m for m in sys.modules if m == "diffdrr" or m.startswith("diffdrr.")]:
        monkeypatch.delitem(sys.modules, name)
    import diffdrr  # noqa: F401 - the alias registers diffdrr.<module>

    yield
    for name in [m for m in sys.modules if m == "diffdrr" or m.startswith("diffdrr.")]:
        sys.modules.pop(name, None)


def _load(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _drr():
    from xvr_b200.data import read, synthetic_ct

    hu, _, affine = synthetic_ct(16)
    return xvr_b200.DRR(read(hu, affine=affine), 1020.0, 16, 8.0, renderer="trilinear", reverse_x_axis=False)


def test_genuine_evaluator_runs_on_the_alias_and_agrees_with_the_mirror(diffdrr_alias):
    """src/xvr/metrics/evaluator.py, unmodified, with our DRR / RigidTransform / DoubleGeodesicSE3 underneath."""
    from xvr_b200.evaluator import Evaluator

    ref = _load("metrics/evaluator.py", "_ref_evaluator")
    drr = _drr()
    g = torch.Generator().manual_seed(0)
    fid = (torch.rand(1, 7, 3, generator=g) - 0.5) * 60.0
    rot = (torch.rand(4, 3, generator=g) - 0.5) * 0.6
    xyz = torch.tensor([0.0, 800.0, 0.0]) + (torch.rand(4, 3, generator=g) - 0.5) * 40.0
    true = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    pred = xvr_b200.convert(rot + 0.03, xyz + 2.0, parameterization="euler_angles", convention="ZXY")
    theirs = ref.Evaluator(drr, fid)(true, pred)
    ours = Evaluator(drr, fid)(true, pred)
    assert torch.allclose(torch.tensor(ours), torch.tensor(theirs), rtol=1e-5, atol=1e-4)
    assert torch.tensor(theirs).shape == (4, 4) and (torch.tensor(theirs) > 0).all()


def test_genuine_sampler_and_loss_run_on_the_alias(diffdrr_alias):
    """src/xvr/model/sampler.py and the geodesic / Dice parts of model/loss.py import and run against the alias."""
    sampler = _load("model/sampler.py", "_ref_sampler")
    torch.manual_seed(0)
    pose = sampler.get_random_pose(-45.0, 45.0, -45.0, 45.0, -15.0, 15.0, -50.0, 50.0, 700.0, 900.0, -50.0, 50.0, 6)
    assert isinstance(pose, xvr_b200.RigidTransform) and pose.matrix.shape == (6, 4, 4)
    torch.manual_seed(0)
    from xvr_b200.sampler import get_random_pose

    mine = get_random_pose(-45.0, 45.0, -45.0, 45.0, -15.0, 15.0, -50.0, 50.0, 700.0, 900.0, -50.0, 50.0, 6)
    assert torch.equal(pose.matrix, mine.matrix)


def test_genuine_initialize_drr_builds_the_module(diffdrr_alias):
    """src/xvr/renderer/load.py: read(volume, mask, labels, orientation, **kw) + the positional DRR constructor."""
    import numpy as np

    from xvr_b200.data import synthetic_ct

    load = _load("renderer/load.py", "_ref_load")
    hu, lab, affine = synthetic_ct(16, with_labels=True)
    drr = load.initialize_drr(hu, lab, "1,2", "PA", 16, 20, 1020.0, 8.0, 7.0, 1.0, -2.0, True, "siddon",
                              read_kwargs={"affine": affine}, drr_kwargs={"voxel_shift": 0.0}, device="cpu")
    d = drr.detector
    assert isinstance(drr, xvr_b200.DRR) and isinstance(drr.renderer, xvr_b200.Siddon)
    assert (d.sdd, d.height, d.width, d.delx, d.dely, d.x0, d.y0, d.reverse_x_axis) == (1020.0, 16, 20, 8.0, 7.0, 1.0, -2.0, True)
    assert drr.renderer.voxel_shift == 0.0 and drr.subject.orientation == "PA"
    assert drr.density.device.type == "cpu" and drr.mask.dtype == torch.uint8
    # labels "1,2" keep only those structures in the density
    keep = torch.isin(lab, torch.tensor([1, 2], dtype=lab.dtype))
    assert (drr.density[~keep] == 0).all() and (drr.density[keep] > 0).any()
    assert np.allclose(np.asarray(drr.subject.volume.get_center()), 0.0, atol=1e-4)  # read() centres the volume


def test_genuine_multiview_consistency_matches_the_mirror(diffdrr_alias):
    """src/xvr/model/loss.py: the pose-pair geodesic (RigidTransform.__getitem__, __matmul__, inverse) and Dice."""
    from xvr_b200.trainer import DiceLoss, PoseRegressionLoss

    ref = _load("model/loss.py", "_ref_loss")
    g = torch.Generator().manual_seed(1)
    mk = lambda: xvr_b200.convert((torch.rand(5, 3, generator=g) - 0.5) * 0.8,  # noqa: E731
                                  torch.tensor([0.0, 800.0, 0.0]) + (torch.rand(5, 3, generator=g) - 0.5) * 50.0,
                                  parameterization="euler_angles", convention="ZXY")
    true, pred = mk(), mk()
    theirs = ref.PoseRegressionLoss(1020.0).multiview_consistency(true, pred)
    ours = PoseRegressionLoss(1020.0).multiview_consistency(true, pred)
    assert theirs.shape == (10,) and torch.allclose(ours, theirs, rtol=1e-5, atol=1e-4)
    a = (torch.rand(3, 4, 8, 8, generator=g) > 0.5).float()
    b = (torch.rand(3, 4, 8, 8, generator=g) > 0.5).float()
    assert torch.allclose(DiceLoss()(a, b), ref.DiceLoss()(a, b))
