"""The genuine registration loop of xvr (``_RegistrarBase.run_test_time_optimization``,
/root/reference/src/xvr/registrar/base.py:198-292) run on the ORACLE renderer and metrics.

    python tests/golden/make_reference_loop_golden.py        # needs /root/reference; writes reference_loop_v1.pt

What this pins: everything the loop itself owns -- torch.optim.Adam(maximize=True) with two parameter groups, the
real torch ReduceLROnPlateau(mode="max"), the order optimizer.step / scheduler.step, the "lr[0] < current_lr"
plateau count (which counts the FIRST iteration), the stop rule, what is logged when.  What it does not pin: the
renderer and the similarity arithmetic, which come from ``oracle`` through adapters (a ``Registration``-shaped
module and an ``imagesim`` lambda) because DiffDRR is not installable here.  The xvr file is loaded unmodified; the
modules it imports are satisfied by name-only stubs, ``.cuda()`` / ``torch.cuda.synchronize`` are patched to no-ops
for the duration of the call (the build container has no GPU).
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
OUT = os.path.join(HERE, "reference_loop_v1.pt")

import oracle  # noqa: E402
from tests.golden.make_golden import scene  # noqa: E402

N_VOL, HEIGHT, DELX, SDD = 48, 40, 1.08821875 * 256.0 / 40, 1020.0
HYPER = dict(lr_rot=1e-2, lr_xyz=1.0, n_itrs=[60], patience=2, threshold=1e-4, max_n_plateaus=3, equalize=False, verbose=0)
ROT0, XYZ0 = [[0.20, -0.10, 0.05]], [[5.0, 800.0, -10.0]]
DROT, DXYZ = [[0.05, -0.04, 0.03]], [[6.0, 9.0, -5.0]]


def load_base():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    pre_spec = importlib.util.spec_from_file_location("_ref_preprocess", os.path.join(REF, "utils/preprocess.py"))
    pre = importlib.util.module_from_spec(pre_spec)
    pre_spec.loader.exec_module(pre)  # the genuine XrayTransforms
    empty = type("Stub", (), {})
    stubs = {
        "matplotlib": mod("matplotlib"), "matplotlib.pyplot": mod("matplotlib.pyplot"),
        "diffdrr": mod("diffdrr"),
        "diffdrr.metrics": mod("diffdrr.metrics", GradientNormalizedCrossCorrelation2d=empty,
                               MultiscaleNormalizedCrossCorrelation2d=empty),
        "diffdrr.registration": mod("diffdrr.registration", Registration=empty),
        "diffdrr.visualization": mod("diffdrr.visualization", plot_drr=None),
        "xvr": mod("xvr", __path__=[]), "xvr.registrar": mod("xvr.registrar", __path__=[]),
        "xvr.renderer": mod("xvr.renderer", initialize_drr=None),
        "xvr.utils": mod("xvr.utils", XrayTransforms=pre.XrayTransforms),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("xvr.registrar.base", os.path.join(REF, "registrar/base.py"))
        base = importlib.util.module_from_spec(spec)
        base.__package__ = "xvr.registrar"
        spec.loader.exec_module(base)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return base


class OracleRegistration(torch.nn.Module):
    """Registration-shaped adapter: Euler ZXY parameters in front of oracle.drr_forward."""

    def __init__(self, density, affinv, rot, xyz):
        super().__init__()
        self.rotation = torch.nn.Parameter(rot.clone())
        self.translation = torch.nn.Parameter(xyz.clone())
        self.density, self.affinv = density, affinv
        det = types.SimpleNamespace(height=HEIGHT, width=HEIGHT)
        self.drr = types.SimpleNamespace(detector=det, rescale_detector_=lambda scale: None)

    @property
    def pose(self):
        return types.SimpleNamespace(convert=lambda p, c: (self.rotation.detach(), self.translation.detach()))

    def forward(self):
        pose = oracle.pose_from_params(self.rotation, self.translation, "euler_angles", "ZXY")
        return oracle.drr_forward(self.density, self.affinv, pose, reorient=oracle.REORIENT["AP"], height=HEIGHT,
                                  width=HEIGHT, delx=DELX, dely=DELX, x0=0.0, y0=0.0, sdd=SDD, reverse_x_axis=False,
                                  renderer="trilinear")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    base = load_base()
    hu, _, affine = scene(N_VOL)
    density = oracle.hu_to_density(hu, 1.0)
    affinv = torch.as_tensor(np.linalg.inv(affine), dtype=torch.float32)[None]
    rot0, xyz0 = torch.tensor(ROT0), torch.tensor(XYZ0)
    with torch.no_grad():
        gt = OracleRegistration(density, affinv, rot0, xyz0)()
    reg = OracleRegistration(density, affinv, rot0 + torch.tensor(DROT), xyz0 + torch.tensor(DXYZ))
    imagesim = lambda x, y: 0.5 * oracle.multiscale_ncc(x, y, (None, 9), (0.5, 0.5)) + 0.5 * oracle.gradient_ncc(x, y, 11, 0.0)  # noqa: E731

    patched = (torch.Tensor.cuda, torch.cuda.synchronize)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        me = types.SimpleNamespace(**HYPER)
        params, nccs, times, alphas = base._RegistrarBase.run_test_time_optimization(me, gt, reg, [1.0], imagesim)
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize = patched
    cases = [("8", 0, 256), ("8,4,2", 0, 256), ("1", 0, 256), ("8", 100, 1436), ("16,8,4,2", 64, 512)]
    scales = {c: base._parse_scales(c[0].split(","), c[1], c[2]) for c in cases}  # base.py:77 splits the string first
    out = {"source": "src/xvr/registrar/base.py:198-292 (run_test_time_optimization), unmodified",
           "parse_scales": scales,
           "scene": dict(n=N_VOL, height=HEIGHT, delx=DELX, sdd=SDD), "hyper": HYPER, "rot0": rot0, "xyz0": xyz0,
           "drot": torch.tensor(DROT), "dxyz": torch.tensor(DXYZ), "gt": gt,
           "params": torch.tensor(params, dtype=torch.float64), "nccs": torch.tensor(nccs, dtype=torch.float64),
           "alphas": torch.tensor(alphas, dtype=torch.float64)}
    torch.save(out, OUT)
    print(f"wrote {OUT}: {len(nccs) - 1} iterations, ncc {nccs[0]:.4f} -> {nccs[-1]:.4f}, final lrs {alphas[-1]}")


if __name__ == "__main__":
    main()
