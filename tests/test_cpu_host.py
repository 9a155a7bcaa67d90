"""CPU suite, part 2: host-side logic of xvr_b200 (no kernels are launched) against the oracle, the C-ABI
export table, and the mirrored conventions."""

import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oracle
import xvr_b200
from oracle import knobs
from xvr_b200 import _conventions as conv
from xvr_b200 import _lib
from xvr_b200.data import read, synthetic_ct, transform_hu_to_density
from xvr_b200.preprocess import XrayTransforms
from xvr_b200.sampler import get_random_pose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_v1.pt"), weights_only=False)


# ------------------------------------------------------------------------------------------------ C-ABI
def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "xvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(xvr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_library_exports_every_declared_symbol():
    if not _lib.LIB.exists():
        pytest.fail(f"{_lib.LIB} is missing: run __graft_entry__.build()")
    handle = ctypes.CDLL(str(_lib.LIB))
    protos = _header_prototypes()
    assert len(protos) >= 19
    for name in protos:
        assert hasattr(handle, name), f"{name} declared in include/xvr_b200.h but not exported"
    assert handle.xvr_abi_version() == 3


def test_bindings_match_header_arity():
    protos = _header_prototypes()
    assert set(protos) == set(_lib.exported_symbols())
    for name, (argtypes, _) in _lib._SIGNATURES.items():
        assert len(argtypes) == protos[name], name


def test_invalid_arguments_return_error_codes_not_crashes():
    lib = _lib.lib()
    assert lib.xvr_reduce_rows(None, 1, 1, None, None) == -1
    assert b"xvr_reduce_rows" in lib.xvr_last_error()
    assert lib.xvr_ncc_fwd(None, None, 1, 1, 8, 8, 0, 1e-5, 1.0, 0, None, None, None, None, None) == -1
    assert lib.xvr_volume_create(0, 4, 4, ctypes.byref(ctypes.c_void_p())) == -1
    # fused registration similarity: sizes are validated (and the workspace sized) without touching the device
    assert lib.xvr_regsim_workspace_floats(1, 8, 8, 9, 11) == -1  # patches larger than the image
    n = lib.xvr_regsim_workspace_floats(1, 256, 256, 9, 11)
    assert n >= 256 * 256 * (1 + 2 + 1 + 2) + 4 * 248 * 248 + 8 * 246 * 246
    assert lib.xvr_regsim(None, None, None, 1, 256, 256, 1e-6, 0.15, 10.0, 9, 11, 0.25, 0.25, 0.5, 1e-5, None, n, None,
                          None, None) == -1
    assert b"xvr_regsim" in lib.xvr_last_error()
    # staged-brick renderer: one or two staging buffers, nothing else
    assert lib.xvr_trilinear_drr_fwd_staged(None, 8, 8, 8, None, None, None, 1, 16, 16, 10, 0, 1e-8, None, None, None,
                                            None) == -1
    assert b"xvr_trilinear_drr_fwd_staged" in lib.xvr_last_error()
    # per-call options: unknown bits and out-of-range fields are invalid arguments, checked before any device work
    assert lib.xvr_siddon_drr_fwd(None, None, 8, 8, 8, None, None, None, 1, 16, 16, 0.5, 1e-8, 0, 3, None, None, 0, None) == -1
    assert b"xvr_siddon_drr_fwd" in lib.xvr_last_error()
    assert lib.xvr_abi_version() == 3


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xvr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_conventions_mirror_oracle_knobs():
    pairs = {
        "COMPOSE_APPLIES_SELF_FIRST": "COMPOSE_APPLIES_SELF_FIRST", "DET_SIGN_S": "DET_SIGN_S", "DET_SIGN_T": "DET_SIGN_T",
        "TRILINEAR_N_POINTS": "TRILINEAR_N_POINTS", "TRILINEAR_STEP": "TRILINEAR_STEP", "RENDER_EPS": "RENDER_EPS",
        "SIDDON_VOXEL_SHIFT_DEFAULT": "SIDDON_VOXEL_SHIFT_DEFAULT", "HU_AIR": "HU_AIR", "HU_BONE": "HU_BONE",
        "NCC_EPS": "NCC_EPS", "GEODESIC_EPS": "GEODESIC_EPS",
        "CONVERT_TRANSLATION_IN_ROTATED_FRAME": "CONVERT_TRANSLATION_IN_ROTATED_FRAME",
    }
    for a, b in pairs.items():
        assert getattr(conv, a) == getattr(knobs, b), a


# ------------------------------------------------------------------------------------------------ pose
@pytest.mark.parametrize("name", list(xvr_b200.N_ANGULAR_COMPONENTS))
def test_convert_matches_oracle_and_golden(name):
    p = GOLD["poses"][name]
    convention = "ZXY" if name == "euler_angles" else None
    T = xvr_b200.convert(p["rot"], p["xyz"], parameterization=name, convention=convention)
    assert torch.allclose(T.matrix, p["matrix"], atol=1e-5)
    R = T.matrix[:, :3, :3]
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand(4, 3, 3), atol=1e-5)
    # convert -> RigidTransform.convert -> convert is the identity on matrices
    rot, xyz = T.convert(name, convention)
    T2 = xvr_b200.convert(rot, xyz, parameterization=name, convention=convention)
    assert torch.allclose(T2.matrix, T.matrix, atol=1e-4)


def test_rigid_transform_algebra():
    g = torch.Generator().manual_seed(0)
    A = xvr_b200.convert(torch.randn(5, 3, generator=g), torch.randn(5, 3, generator=g) * 100,
                         parameterization="euler_angles", convention="ZXY")
    Bt = xvr_b200.convert(torch.randn(1, 3, generator=g), torch.randn(1, 3, generator=g) * 100,
                          parameterization="axis_angle")
    pts = torch.randn(5, 7, 3, generator=g)
    assert torch.allclose(A.compose(Bt)(pts), Bt(A(pts)), atol=1e-3)  # A first, then B; batch of 1 broadcasts
    assert torch.allclose(A.inverse()(A(pts)), pts, atol=1e-3)
    assert torch.allclose((A @ A.inverse()).matrix, torch.eye(4).expand(5, 4, 4), atol=1e-4)
    assert len(A) == 5 and len(A[torch.tensor([True, False, True, False, False])]) == 2
    assert torch.equal(A[1:3].matrix, A.matrix[1:3])
    assert torch.allclose(A.compose(Bt).matrix, oracle.compose(A.matrix, Bt.matrix))
    assert torch.allclose(A.inverse().matrix, oracle.invert(A.matrix), atol=1e-5)


def test_degrees_and_sampler():
    torch.manual_seed(0)
    pose = get_random_pose(-45, 45, -45, 45, -15, 15, -50, 50, 700, 900, -50, 50, 16)
    rot, xyz = pose.convert("euler_angles", "ZXY", degrees=True)
    assert rot[:, :2].abs().max() <= 45 + 1e-3 and rot[:, 2].abs().max() <= 15 + 1e-3
    assert (xyz[:, 1] >= 700 - 1e-2).all() and (xyz[:, 1] <= 900 + 1e-2).all()
    # the camera orbits the isocenter: |source| == |xyz|
    assert torch.allclose(pose.matrix[:, :3, 3].norm(dim=-1), xyz.norm(dim=-1), rtol=1e-5)


def test_geodesic_against_oracle():
    from xvr_b200.metrics import DoubleGeodesicSE3

    g = torch.Generator().manual_seed(2)
    A = xvr_b200.convert(torch.randn(6, 3, generator=g) * 0.5, torch.randn(6, 3, generator=g) * 30,
                         parameterization="euler_angles", convention="ZXY")
    B = xvr_b200.convert(torch.randn(6, 3, generator=g) * 0.5, torch.randn(6, 3, generator=g) * 30,
                         parameterization="euler_angles", convention="ZXY")
    for a, b in zip(DoubleGeodesicSE3(1020.0)(A, B), oracle.double_geodesic(A.matrix, B.matrix, 1020.0)):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ DRR module
def _drr(renderer="trilinear", **kw):
    sub = read(GOLD["hu"], GOLD["labels"], affine=GOLD["affine"].numpy(), center_volume=False)
    d = GOLD["detector"]
    return xvr_b200.DRR(sub, d["sdd"], d["height"], d["delx"], d["width"], d["dely"], d["x0"], d["y0"],
                        reverse_x_axis=d["reverse_x_axis"], renderer=renderer, **kw)


def test_detector_matches_oracle():
    drr = _drr()
    d = GOLD["detector"]
    pose = xvr_b200.convert(GOLD["rot"], GOLD["xyz"], parameterization="euler_angles", convention="ZXY")
    src, tgt = drr.detector(pose, None)
    osrc, otgt = oracle.detector_rays(pose.matrix, oracle.REORIENT["AP"], d["height"], d["width"], d["delx"],
                                      d["dely"], d["x0"], d["y0"], d["sdd"], d["reverse_x_axis"])
    assert torch.allclose(src, osrc, atol=1e-4) and torch.allclose(tgt, otgt, atol=1e-3)
    # the in-kernel pixel basis reproduces the materialised grid
    o, u, v = (torch.tensor(t) for t in drr.detector.pixel_basis())
    i, j = torch.meshgrid(torch.arange(d["height"]), torch.arange(d["width"]), indexing="ij")
    grid = o + i[..., None] * u + j[..., None] * v
    assert torch.allclose(grid.reshape(1, -1, 3), drr.detector.target, atol=1e-4)
    assert torch.equal(drr.density, GOLD["density"])


def test_drr_surface_used_by_xvr():
    drr = _drr("siddon", voxel_shift=0.0)
    assert drr.renderer.voxel_shift == 0.0
    assert isinstance(copy.deepcopy(drr), xvr_b200.DRR)
    drr.set_intrinsics_(sdd=900.0, height=40, width=30, delx=2.0, dely=2.5, x0=-1.0, y0=2.0)
    assert (drr.detector.height, drr.detector.width, drr.detector.sdd) == (40, 30, 900.0)
    assert drr.detector.target.shape == (1, 1200, 3) and drr.renderer.detector_hw == (40, 30)
    drr.rescale_detector_(0.25)
    assert (drr.detector.height, drr.detector.width) == (10, 7) and drr.detector.delx == 8.0
    img = torch.zeros(2, 1, 70)
    assert drr.reshape_transform(img, batch_size=2).shape == (2, 1, 10, 7)
    # mutations of model/utils.py:162-171
    drr.density = None
    drr.register_buffer("volume", GOLD["hu"])
    drr.register_buffer("center", torch.zeros(1, 3))
    assert hasattr(drr, "mask") and drr.affine_inverse(torch.zeros(1, 1, 3)).shape == (1, 1, 3)
    with pytest.raises(RuntimeError):
        drr(xvr_b200.convert(GOLD["rot"], GOLD["xyz"], parameterization="euler_angles", convention="ZXY"))


def test_no_cpu_fallback():
    drr = _drr()
    pose = xvr_b200.convert(GOLD["rot"], GOLD["xyz"], parameterization="euler_angles", convention="ZXY")
    with pytest.raises(_lib.XvrB200Error):
        drr(pose)
    from xvr_b200.metrics import NormalizedCrossCorrelation2d

    with pytest.raises(_lib.XvrB200Error):
        NormalizedCrossCorrelation2d()(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


def test_projection_round_trip():
    drr = _drr()
    pose = xvr_b200.convert(GOLD["rot"], GOLD["xyz"], parameterization="euler_angles", convention="ZXY")
    pts = torch.randn(3, 5, 3) * 20
    uv = drr.perspective_projection(pose, pts)
    back = drr.inverse_projection(pose, uv)
    # the back-projected point lies on the source -> point ray
    src = pose.matrix[:, None, :3, 3]
    a, b = pts - src, back - src
    cos = (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))
    assert torch.allclose(cos, torch.ones_like(cos), atol=1e-5)


def test_registration_module():
    drr = _drr()
    reg = xvr_b200.Registration(drr, GOLD["rot"][:1], GOLD["xyz"][:1], "euler_angles", "ZXY")
    assert {n for n, _ in reg.named_parameters()} == {"rotation", "translation"}
    assert torch.allclose(reg.pose.matrix, xvr_b200.convert(GOLD["rot"][:1], GOLD["xyz"][:1],
                                                            parameterization="euler_angles", convention="ZXY").matrix)


# ------------------------------------------------------------------------------------------------ data / transforms
def test_hu_to_density_matches_oracle():
    hu, _, _ = synthetic_ct(24, seed=3)
    for m in (1.0, 4.5):
        assert torch.allclose(transform_hu_to_density(hu, m), oracle.hu_to_density(hu, m), atol=1e-7)


def test_xray_transforms_match_oracle():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 1, 40, 40, generator=g) * 50
    assert torch.allclose(XrayTransforms(40)(x), oracle.xray_transforms(x, 40), atol=1e-6)
    assert torch.allclose(XrayTransforms(20, 16)(x), oracle.xray_transforms(x, 20, 16), atol=1e-6)
    y = XrayTransforms(40, equalize=True)(x)
    assert y.shape == x.shape and torch.isfinite(y).all()


def test_read_centres_the_volume():
    hu, lab, affine = synthetic_ct(16, with_labels=True)
    sub = read(hu, lab, labels=[2, 3], affine=affine)
    assert np.allclose(sub.volume.get_center(), 0.0)
    assert sub.density.shape == (16, 16, 16) and float(sub.density.max()) <= 1.0
    assert float(sub.density[lab < 2].abs().max()) == 0.0


def test_evaluator_metrics():
    """xvr's Evaluator (metrics/evaluator.py): zero for identical poses; a pure in-plane shift of t mm gives
    mTRE = t, mPE = t * sdd / depth (magnification) and a positive geodesic."""
    from xvr_b200.evaluator import Evaluator

    drr = _drr()
    fid = torch.tensor([[[0.0, 0.0, 0.0], [10.0, -5.0, 8.0], [-12.0, 6.0, -7.0]]])
    rot, xyz = torch.zeros(1, 3), torch.tensor([[0.0, 800.0, 0.0]])
    true = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    assert Evaluator(drr, fid)(true, true) == pytest.approx([0.0, 0.0, 0.0, 0.0], abs=1e-3)
    pred = xvr_b200.convert(rot, xyz + torch.tensor([[3.0, 0.0, 0.0]]), parameterization="euler_angles", convention="ZXY")
    mpe, mrpe, mtre, dgeo = Evaluator(drr, fid)(true, pred)
    assert mtre == pytest.approx(3.0, abs=1e-3)
    assert 3.0 < mpe < 3.0 * 1020.0 / 700.0  # magnified by sdd / depth, depth in (700, 900) for these fiducials
    assert mrpe > 0 and dgeo == pytest.approx(3.0, abs=1e-3)


def test_compat_package_serves_every_diffdrr_import_of_xvr(monkeypatch):
    """Every ``from diffdrr.<module> import <name>`` statement in xvr's sources (listed from /root/reference/src and
    /root/reference/scripts) resolves through the alias package xvr_b200/compat/diffdrr."""
    import importlib
    import os
    import sys

    import xvr_b200.compat

    monkeypatch.syspath_prepend(os.path.dirname(xvr_b200.compat.__file__))
    for name in [m for m in sys.modules if m == "diffdrr" or m.startswith("diffdrr.")]:
        monkeypatch.delitem(sys.modules, name)
    wanted = {
        "diffdrr.drr": ["DRR"],
        "diffdrr.pose": ["RigidTransform", "convert", "make_matrix"],
        "diffdrr.registration": ["Registration", "N_ANGULAR_COMPONENTS"],
        "diffdrr.metrics": ["DoubleGeodesicSE3", "MultiscaleNormalizedCrossCorrelation2d",
                            "GradientNormalizedCrossCorrelation2d"],
        "diffdrr.data": ["read", "load_example_ct", "transform_hu_to_density"],
        "diffdrr.utils": ["resample"],
        "diffdrr.visualization": ["plot_drr", "plot_mask"],
    }
    for module, names in wanted.items():
        mod = importlib.import_module(module)
        for n in names:
            assert hasattr(mod, n), f"{module}.{n}"
    assert "xvr_b200" in importlib.import_module("diffdrr").__version__


def test_resample_identity_and_shapes():
    from xvr_b200.utils import resample

    img = torch.rand(2, 1, 40, 36)
    assert torch.equal(resample(img, 1020.0, 0.2), img)
    for kw in (dict(new_focal_len=1200.0), dict(new_delx=0.3), dict(new_x0=4.0, new_y0=-2.0), dict(new_delx=0.15)):
        out = resample(img, 1020.0, 0.2, 0.0, 0.0, **kw)
        assert out.shape == img.shape and torch.isfinite(out).all()
    # doubling the pixel size shows the old image, shrunk, in the centre: the border is padding
    big = resample(torch.ones(1, 1, 40, 40), 1020.0, 0.2, new_delx=0.4)
    assert big[0, 0, 20, 20] > 0.99 and big[0, 0, 2, 2] == 0


@pytest.mark.parametrize("renderer", ["trilinear", "siddon"])
def test_empty_batches_return_empty_results_without_a_launch(monkeypatch, renderer):
    """Empty in -> empty out on every renderer / similarity entry, forward and backward (xvr indexes its batch with a
    `keep` mask that can be all-False, /root/reference/src/xvr/model/trainer.py:202-204).  Host logic only: the
    device check is lifted and any C-ABI call fails the test."""
    from tests._scene import pixel_size
    from xvr_b200 import drr as drr_mod
    from xvr_b200 import metrics, renderers

    def no_launch(name, *args):
        raise AssertionError(f"{name} launched for an empty batch")

    for mod in (renderers, drr_mod, metrics):
        monkeypatch.setattr(mod, "cuda_f32", lambda t, what: t.to(torch.float32).contiguous())
        monkeypatch.setattr(mod, "call", no_launch)
        monkeypatch.setattr(mod, "stream", lambda: None)
    monkeypatch.setattr(renderers._VolumeTexture, "get", lambda self, volume: None)

    hu, _, affine = synthetic_ct(16)
    drr = xvr_b200.DRR(read(hu, affine=affine), 1020.0, 8, pixel_size(8), renderer=renderer, reverse_x_axis=False)
    rot = torch.zeros(0, 3, requires_grad=True)
    xyz = torch.zeros(0, 3, requires_grad=True)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    img = drr(pose)
    assert img.shape == (0, 1, 8, 8)
    img.sum().backward()
    assert rot.grad.shape == (0, 3) and xyz.grad.shape == (0, 3)
    assert drr(rot, xyz, parameterization="euler_angles", convention="ZXY").shape == (0, 1, 8, 8)
    with torch.no_grad():
        source, target = drr.detector(pose, None)
        raylen = (target - source).norm(dim=-1).unsqueeze(1)
        assert drr.renderer(drr.density, drr.affine_inverse(source), drr.affine_inverse(target),
                            raylen).shape == (0, 1, 64)
        assert drr.renderer(drr.density, torch.zeros(2, 1, 3), torch.zeros(2, 0, 3),
                            torch.zeros(2, 1, 0)).shape == (2, 1, 0)

    x = torch.zeros(0, 1, 32, 32, requires_grad=True)
    y = torch.zeros(0, 1, 32, 32)
    s1 = metrics.MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])(x, y)
    s2 = metrics.GradientNormalizedCrossCorrelation2d(11, sigma=0.0)(x, y)
    assert s1.shape == s2.shape == (0,)
    (s1.sum() + s2.sum()).backward()
    assert x.grad.shape == x.shape


def test_checkpoint_round_trip_and_pose_prediction(tmp_path):
    """*.pth files carry the reference's keys (trainer.py:318-332), load_model rebuilds the regressor from the
    stored config (network.py:57-78) and predict_pose resamples / crops / normalises as inference.py:9-39."""
    from xvr_b200.inference import construct_antipode, correct_pose, load_model, predict_pose, save_checkpoint
    from xvr_b200.trainer import PoseRegressor, WarmupCosineSchedule

    torch.manual_seed(0)
    model = PoseRegressor("resnet18", "quaternion_adjugate", "ZXY", height=64, unit_conversion_factor=1000.0).eval()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = WarmupCosineSchedule(opt, 10, 100)
    config = dict(model_name="resnet18", parameterization="quaternion_adjugate", convention="ZXY",
                  norm_layer="groupnorm", height=64, sdd=1020.0, delx=4.0, unit_conversion_factor=1000.0)
    path = save_checkpoint(tmp_path / "0000.pth", model, opt, sched, itr=7, model_number=0, config=config)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "itr", "model_number",
                         "date", "config"}
    assert ckpt["itr"] == 7 and ckpt["config"]["height"] == 64
    # timm's ResNet parameter names (what an xvr-written model_state_dict holds)
    keys = set(ckpt["model_state_dict"])
    assert {"backbone.conv1.weight", "backbone.bn1.weight", "backbone.layer2.0.downsample.0.weight",
            "backbone.layer4.1.bn2.bias", "xyz_regression.weight", "rot_regression.bias"} <= keys
    assert not any(k.startswith("backbone.fc") for k in keys)

    loaded, cfg, date = load_model(path, meta=True, device="cpu")
    assert not loaded.training and cfg == config and date is not None
    img = torch.rand(2, 1, 120, 150)
    pose_a, seen = predict_pose(model, config, img, 1020.0, 2.0, 2.0, 0.0, 0.0)
    pose_b, _ = predict_pose(loaded, cfg, img, 1020.0, 2.0, 2.0, 0.0, 0.0)
    assert seen.shape == (2, 1, 64, 64)
    assert torch.equal(pose_a.matrix, pose_b.matrix)
    with pytest.raises(AssertionError):
        predict_pose(model, config, img, 1020.0, 2.0, 2.5, 0.0, 0.0)

    # models saved before the unit switch have no unit_conversion_factor: the reference falls back to 1.0
    old = {k: v for k, v in config.items() if k != "unit_conversion_factor"}
    save_checkpoint(tmp_path / "old.pth", model, opt, sched, 0, 0, old)
    assert load_model(tmp_path / "old.pth", device="cpu")[0].unit_conversion_factor == 1.0

    anti = construct_antipode(pose_a)
    r0, _ = pose_a.convert("euler_angles", "ZXY")
    r1, _ = anti.convert("euler_angles", "ZXY")
    assert torch.allclose(r1[:, 1:], r0[:, 1:] * torch.tensor([-1.0, 1.0]), atol=1e-4)
    assert torch.allclose(construct_antipode(anti).matrix, pose_a.matrix, atol=5e-3)
    assert correct_pose(pose_a, None) is pose_a
    shift = xvr_b200.convert(torch.zeros(1, 3), torch.tensor([[1.0, 2.0, 3.0]]), parameterization="euler_angles",
                             convention="ZXY")
    assert torch.allclose(correct_pose(pose_a, shift.matrix[0]).matrix, pose_a.compose(shift).matrix)


def test_fused_registration_similarity_host_wiring(monkeypatch):
    """Host side of xvr_regsim: weights, transform constants and workspace size reach the C-ABI as documented in
    include/xvr_b200.h; the Registrar only selects it for the configuration it covers.  No kernel is launched."""
    from xvr_b200 import metrics
    from xvr_b200.registrar import Registrar

    seen = {}

    def fake_call(name, *args):
        seen[name] = args

    monkeypatch.setattr(metrics, "cuda_f32", lambda t, what: t.to(torch.float32).contiguous())
    monkeypatch.setattr(metrics, "call", fake_call)
    monkeypatch.setattr(metrics, "stream", lambda: None)
    fixed = torch.rand(2, 1, 40, 48)
    sim = metrics.RegistrationSimilarity(fixed, mncc_patch_size=9, gncc_patch_size=11, beta=0.3)
    assert "xvr_sobel_fwd" in seen and sim.fixed_sobel.shape == (2, 2, 40, 48)
    moving = torch.rand(2, 1, 40, 48, requires_grad=True)
    score = sim(moving)
    args = seen["xvr_regsim"]
    B, H, W, std_eps, mean, inv_std, p, q, w_global, w_patch, w_grad, eps, _, n_work = args[3:17]
    assert (B, H, W, p, q) == (2, 40, 48, 9, 11)
    assert std_eps == pytest.approx(1e-6) and mean == pytest.approx(0.15) and inv_std == pytest.approx(10.0)
    assert (w_global, w_patch, w_grad) == pytest.approx((0.15, 0.15, 0.7)) and eps == pytest.approx(conv.NCC_EPS)
    assert n_work == _lib.lib().xvr_regsim_workspace_floats(2, 40, 48, 9, 11) > 0
    assert score.shape == ()
    (3.0 * score).backward()  # backward = saved gradient x upstream scalar (zeros here: nothing was launched)
    assert moving.grad.shape == moving.shape
    with pytest.raises(_lib.XvrB200Error):
        metrics.RegistrationSimilarity(torch.rand(1, 1, 8, 8), 9, 11)(torch.rand(1, 1, 8, 8))

    drr = object()
    assert Registrar(drr, fused_similarity=True).fused_similarity
    assert not Registrar(drr, fused_similarity=True, equalize=True).fused_similarity
    assert not Registrar(drr, fused_similarity=True, sigma=1.0).fused_similarity
    assert Registrar(drr).fused_similarity and not Registrar(drr, fused_similarity=False).fused_similarity


# ------------------------------------------------------------------------------------------------ round-2 additions
def test_label_cache_is_keyed_by_tensor_identity_not_by_address():
    """The reference's Trainer.load builds a fresh float mask per iteration; the allocator reuses the block, so
    pointer + shape + version 0 repeat for a DIFFERENT label map.  The cache must not serve the old entry."""
    from xvr_b200.renderers import _LabelCache

    cache = _LabelCache()
    a = torch.zeros(4, 4, 4)
    a[0, 0, 0] = 3.0
    lab_a, c_a = cache.get(a)
    assert c_a == 4 and cache.get(a)[0] is lab_a  # same object: a hit
    key = next(iter(cache.entries))
    del a
    b = torch.zeros(4, 4, 4)
    b[1, 1, 1] = 7.0
    # force the collision the allocator produces on the GPU: file b under a's key
    entry = cache.entries.pop(key)
    cache.entries[(b.data_ptr(), b._version, tuple(b.shape), b.dtype, b.device)] = entry
    lab_b, c_b = cache.get(b)
    assert c_b == 8 and lab_b[1, 1, 1] == 7 and lab_b is not lab_a


def test_registrar_save_writes_the_reference_schema(tmp_path):
    """parameters.pt carries the keys of the reference's _RegistrarBase.save (registrar/base.py:355-394) and
    round-trips; B != 1 is rejected before any device work."""
    from tests._scene import make_subject
    from xvr_b200.registrar import Registrar

    drr = xvr_b200.DRR(make_subject(16), 1020.0, 8, 2.0, renderer="trilinear")
    reg = Registrar(drr, scales="4,2", n_itrs="50,20", provenance={"volume": "ct.nii.gz", "labels": [1, 2]})
    info = {"params": [[0.0] * 6, [0.1] * 6], "nccs": [0.5, 0.6], "times": [0.0, 0.01], "alphas": [[1e-2, 1.0]] * 2}
    traj = reg.trajectory(info)
    assert list(traj.columns if hasattr(traj, "columns") else traj) == list(Registrar.COLUMNS)
    intr = dict(sdd=1020.0, height=8, width=8, delx=2.0, dely=2.0, x0=-1.5, y0=0.5)
    eye = torch.eye(4)[None]
    reg.save(tmp_path, torch.zeros(1, 1, 8, 8), None, None, "xray_0001", intr, eye, eye * 2,
             dict(pf_to_af=None, runtime=0.01, trajectory=traj))
    got = torch.load(tmp_path / "parameters.pt", weights_only=False)
    assert set(got) == {"drr", "xray", "optimization", "init_pose", "final_pose", "pf_to_af", "runtime", "trajectory"}
    assert set(got["drr"]) == {"volume", "mask", "labels", "orientation", "sdd", "height", "width", "delx", "dely",
                               "x0", "y0", "reverse_x_axis", "renderer", "read_kwargs", "drr_kwargs"}
    assert set(got["xray"]) == {"filename", "crop", "subtract_background", "linearize", "reducefn"}
    assert set(got["optimization"]) == {"equalize", "init_only", "scales", "n_itrs", "parameterization", "convention",
                                        "lr_rot", "lr_xyz", "patience", "max_n_plateaus"}
    assert got["drr"]["volume"].name == "ct.nii.gz" and got["drr"]["renderer"] == "trilinear"
    assert got["optimization"]["scales"] == "4,2" and got["optimization"]["n_itrs"] == "50,20"
    assert torch.equal(got["final_pose"], eye * 2) and got["drr"]["x0"] == -1.5
    with pytest.raises(ValueError, match="ONE X-ray"):
        reg.run(torch.zeros(2, 1, 8, 8), xvr_b200.RigidTransform(torch.eye(4).repeat(2, 1, 1)))


def test_reference_private_correct_pose_signature():
    """inference._correct_pose(pose, warp, volume, invert) is callable the way the reference calls it
    (/root/reference/src/xvr/model/inference.py:42-48)."""
    from xvr_b200 import inference

    pose = xvr_b200.convert(torch.tensor([[0.1, 0.2, 0.3]]), torch.tensor([[1.0, 2.0, 3.0]]),
                            parameterization="euler_angles", convention="ZXY")
    assert inference._correct_pose(pose, None, None, False) is pose
    shift = xvr_b200.convert(torch.tensor([[0.0, 0.1, 0.0]]), torch.tensor([[5.0, 0.0, 0.0]]),
                             parameterization="euler_angles", convention="ZXY")
    fwd = inference._correct_pose(pose, shift.matrix[0], None, False)
    back = inference._correct_pose(fwd, shift.matrix[0], None, True)
    assert torch.allclose(fwd.matrix, pose.compose(shift).matrix)
    assert torch.allclose(back.matrix, pose.matrix, atol=1e-5)
    with pytest.raises(NotImplementedError):
        inference._correct_pose(pose, "warp.mat", "ct.nii.gz", False)


def test_options_word_encoding():
    """The per-call options word of include/xvr_b200.h, as the host composes it."""
    from xvr_b200._lib import options, opts_word

    assert opts_word() == 0
    with options(ksplit=2, siddon_walk=True, volgrad="gather", siddon_tol="exact"):
        assert opts_word() == (3 | 0x10 | 0x20 | (1 << 8))
    assert opts_word() == 0
    with pytest.raises(TypeError):
        options(nonsense=1)
    with pytest.raises(ValueError), options(ksplit=7):
        pass
    assert opts_word() == 0


def test_label_brick_table_matches_brute_force():
    """renderers._labels_with_brick_table: label bytes first, then (256-byte aligned) the per-brick table of
    XVR_OPT_LABEL_BRICKS -- the label shared by every voxel of the 8^3 brick grown by one voxel (outside = 0), else 255."""
    from xvr_b200.renderers import LABEL_BRICK, _labels_with_brick_table

    g = torch.Generator().manual_seed(5)
    uniform = mixed = 0
    for shape in [(40, 32, 48), (19, 8, 27), (7, 33, 10)]:
        lab = torch.zeros(shape, dtype=torch.uint8)
        lab[2:, 1:, 3:] = 1  # a body that reaches three faces of the volume
        lab[5:9, 2:6, 4:20] = 3
        lab[tuple(int(torch.randint(0, n, (1,), generator=g)) for n in shape)] = 254
        view = _labels_with_brick_table(lab.float())
        n = lab.numel()
        assert torch.equal(view, lab) and view.dtype == torch.uint8
        buf = torch.empty(0, dtype=torch.uint8).set_(view.untyped_storage())
        nb = [(d + LABEL_BRICK - 1) // LABEL_BRICK for d in shape]
        table = buf[(n + 255) // 256 * 256:].view(nb)
        padded = torch.zeros([b * LABEL_BRICK + 2 for b in nb], dtype=torch.uint8)
        padded[1:shape[0] + 1, 1:shape[1] + 1, 1:shape[2] + 1] = lab
        for b0 in range(nb[0]):
            for b1 in range(nb[1]):
                for b2 in range(nb[2]):
                    blk = padded[b0 * 8:b0 * 8 + 10, b1 * 8:b1 * 8 + 10, b2 * 8:b2 * 8 + 10]
                    want = int(blk.flatten()[0]) if bool((blk == blk.flatten()[0]).all()) else 255
                    assert int(table[b0, b1, b2]) == want, (shape, b0, b1, b2)
        uniform += int((table != 255).sum())
        mixed += int((table == 255).sum())
    assert uniform > 0 and mixed > 0
