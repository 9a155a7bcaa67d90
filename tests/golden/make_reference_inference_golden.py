"""Golden run of the GENUINE ``src/xvr/model/inference.py`` (unmodified, loaded by path) for the xvr-owned arithmetic
in front of a model-based registration: which intrinsics ``_resample_xray`` asks ``resample`` for, what
``predict_pose`` feeds the network (resample -> centre crop -> XrayTransforms), and the Euler-angle arithmetic of
``_construct_antipode``.

    python tests/golden/make_reference_inference_golden.py      # needs /root/reference; writes reference_inference_v1.pt

DiffDRR is not installable here, so ``diffdrr`` is a NAME-ONLY stub: ``resample`` records its arguments and hands the
image back, ``convert`` hands its arguments back, ``RigidTransform`` only carries the (rot, xyz) its ``convert``
returns.  ``..utils`` is served by the genuine ``utils/preprocess.py`` (``XrayTransforms``) and a dummy ``get_4x4``
(file IO).  ``Tensor.cuda`` is the identity for the duration.  What is recorded is therefore xvr's own computation.
"""

import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference/src/xvr"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_inference_v1.pt")


def load(relpath, name, package=None):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    calls = []

    class RigidTransform:  # name-only stand-in
        def __init__(self, rot, xyz):
            self.rot, self.xyz = rot, xyz

        def convert(self, parameterization, convention):
            calls.append(("RigidTransform.convert", parameterization, convention))
            return self.rot.clone(), self.xyz.clone()

    def convert(rot, xyz, parameterization=None, convention=None):
        return {"rot": rot, "xyz": xyz, "parameterization": parameterization, "convention": convention}

    def resample(img, *args):
        calls.append(("resample", tuple(float(a) for a in args)))
        return img

    diffdrr = types.ModuleType("diffdrr")
    pose = types.ModuleType("diffdrr.pose")
    pose.RigidTransform, pose.convert = RigidTransform, convert
    utils_d = types.ModuleType("diffdrr.utils")
    utils_d.resample = resample
    sys.modules.update({"diffdrr": diffdrr, "diffdrr.pose": pose, "diffdrr.utils": utils_d})

    pre = load("utils/preprocess.py", "_refpkg.utils.preprocess")
    pkg = types.ModuleType("_refpkg")
    pkg.__path__ = []
    utils_x = types.ModuleType("_refpkg.utils")
    utils_x.XrayTransforms = pre.XrayTransforms
    utils_x.get_4x4 = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("file IO is out of scope"))
    model_pkg = types.ModuleType("_refpkg.model")
    model_pkg.__path__ = []
    sys.modules.update({"_refpkg": pkg, "_refpkg.utils": utils_x, "_refpkg.model": model_pkg})
    inf = load("model/inference.py", "_refpkg.model.inference", package="_refpkg.model")

    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        g = torch.Generator().manual_seed(11)
        out = {"source": "src/xvr/model/inference.py", "torch": torch.__version__, "cases": []}
        config = {"height": 64, "delx": 2.5, "sdd": 1020.0}
        seen = {}

        def model(x):
            seen["x"] = x.clone()
            return "pose"

        for shape, (sdd, delx, x0, y0) in (((2, 1, 120, 150), (990.0, 0.8, 3.0, -2.0)), ((1, 1, 200, 96), (1100.0, 0.5, 0.0, 0.0)),
                                           ((1, 1, 64, 64), (1020.0, 2.5, 0.0, 0.0))):
            img = torch.rand(*shape, generator=g) * 3.0
            calls.clear()
            pose_out, seen_img = inf.predict_pose(model, config, img, sdd, delx, delx, x0, y0)
            assert pose_out == "pose"
            out["cases"].append({"img": img, "intrinsics": (sdd, delx, x0, y0), "config": dict(config),
                                 "resample_args": [c[1] for c in calls if c[0] == "resample"][0],
                                 "model_input": seen["x"], "returned_img": seen_img})

        rot = (torch.rand(5, 3, generator=g) - 0.5) * 3.0
        xyz = torch.rand(5, 3, generator=g) * 100.0
        calls.clear()
        anti = inf._construct_antipode(RigidTransform(rot, xyz))
        out["antipode"] = {"rot": rot, "xyz": xyz, "rot_out": anti["rot"], "xyz_out": anti["xyz"],
                           "convert_args": (anti["parameterization"], anti["convention"]), "pose_convert": calls[0][1:]}
        out["correct_pose_without_warp_is_identity"] = inf._correct_pose("p", None, None, False) == "p"
    finally:
        torch.Tensor.cuda = cuda
    torch.save(out, OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
