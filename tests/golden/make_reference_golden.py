"""Golden vectors produced by the GENUINE reference code (not by the oracle).

    python tests/golden/make_reference_golden.py        # needs /root/reference; writes reference_v1.pt

Only two files on xvr's hot path import without DiffDRR and can therefore run in the build container:
``src/xvr/utils/preprocess.py`` (XrayTransforms / Standardize / Equalize) and
``src/xvr/model/scheduler.py`` (WarmupCosineSchedule, used by Trainer).  They are loaded by path, unmodified, and
their outputs on seeded inputs are stored together with the inputs.  Two more files contain xvr-OWNED arithmetic
behind a ``diffdrr`` import: ``model/sampler.py`` (the order and ranges of the random pose draws) and
``model/loss.py`` (DiceLoss / DiceMetric).  They are loaded with a stub ``diffdrr`` in ``sys.modules`` that only
supplies the imported NAMES (``convert`` hands its arguments back; the metric classes are empty): what is recorded
is the genuine xvr code's own computation, nothing of DiffDRR is emulated.  These vectors PIN the corresponding pieces of
the oracle and of the product (tests/test_cpu_reference_golden.py); everything that needs DiffDRR itself remains
"parity unpinned" (DESIGN.md section 3).
"""

import importlib.util
import os
import sys

import torch

REF = "/root/reference/src/xvr"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_v1.pt")


def load(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    pre = load("utils/preprocess.py", "_ref_preprocess")
    sch = load("model/scheduler.py", "_ref_scheduler")
    g = torch.Generator().manual_seed(7)
    out = {"source": {"preprocess": "src/xvr/utils/preprocess.py", "scheduler": "src/xvr/model/scheduler.py"},
           "torch": torch.__version__}

    # ---- XrayTransforms: DRR-like non-negative images, batch-global Standardize, bilinear Resize, Normalize
    x = torch.rand(3, 1, 40, 36, generator=g) ** 2 * 7.5
    x[1] *= 0.3
    cases = {}
    for name, kw in {"same_size": dict(height=40, width=36), "down": dict(height=24), "up_eq": dict(height=48, width=44, equalize=True),
                     "custom_norm": dict(height=32, mean=0.3, std=0.25)}.items():
        cases[name] = {"kwargs": kw, "out": pre.XrayTransforms(**kw)(x)}
    out["xray_transforms"] = {"x": x, "cases": cases}
    out["standardize"] = {"x": x, "out": pre.Standardize()(x), "out_eps": pre.Standardize(eps=1e-3)(x)}
    out["equalize"] = {"x": pre.Standardize()(x), "out": pre.Equalize()(pre.Standardize()(x)),
                       "out_coarse": pre.Equalize(n_bins=32, tau=0.05)(pre.Standardize()(x))}

    # ---- WarmupCosineSchedule / IdentitySchedule: learning rate after every optimiser step
    def lrs(make, steps):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=2e-4)
        s = make(opt)
        seq = [s.get_last_lr()[0]]
        for _ in range(steps):
            opt.step()
            s.step()
            seq.append(s.get_last_lr()[0])
        return torch.tensor(seq, dtype=torch.float64)

    out["schedule"] = {
        "warmup_cosine_250_of_2500": lrs(lambda o: sch.WarmupCosineSchedule(o, 250, 2500), 300),
        "warmup_cosine_fractional": lrs(lambda o: sch.WarmupCosineSchedule(o, 2.5, 40.0), 45),
        "identity": lrs(lambda o: sch.IdentitySchedule(o), 5),
    }
    # ---- xvr-owned logic behind a diffdrr import: load with name-only stubs
    import types

    stub = types.ModuleType("diffdrr")
    stub.pose = types.ModuleType("diffdrr.pose")
    stub.pose.convert = lambda rot, xyz, **kw: (rot, xyz, kw)
    stub.metrics = types.ModuleType("diffdrr.metrics")
    stub.metrics.DoubleGeodesicSE3 = type("DoubleGeodesicSE3", (torch.nn.Module,), {})
    stub.metrics.MultiscaleNormalizedCrossCorrelation2d = type("MultiscaleNormalizedCrossCorrelation2d", (torch.nn.Module,), {})
    saved = {k: sys.modules.get(k) for k in ("diffdrr", "diffdrr.pose", "diffdrr.metrics")}
    sys.modules.update({"diffdrr": stub, "diffdrr.pose": stub.pose, "diffdrr.metrics": stub.metrics})
    try:
        smp = load("model/sampler.py", "_ref_sampler")
        los = load("model/loss.py", "_ref_loss")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ranges = dict(alphamin=-45.0, alphamax=45.0, betamin=-170.0, betamax=190.0, gammamin=-15.0, gammamax=15.0,
                  txmin=-50.0, txmax=50.0, tymin=700.0, tymax=900.0, tzmin=-50.0, tzmax=50.0)
    torch.manual_seed(123)
    rot, xyz, kw = smp.get_random_pose(**ranges, batch_size=7)
    out["sampler"] = {"ranges": ranges, "seed": 123, "batch_size": 7, "rot": rot, "xyz": xyz, "convert_kwargs": kw}
    a = torch.rand(3, 4, 12, 10, generator=g) > 0.6
    b = torch.rand(3, 4, 12, 10, generator=g) > 0.5
    a[1, 2] = False
    b[1, 2] = False  # a channel empty in both masks: nan in the metric, ignored by nanmean
    a[2, 1:] = False
    b[2, 1:] = False  # every foreground channel empty: nanmean is nan, nan_to_num -> 0
    out["dice"] = {"a": a, "b": b, "loss": los.DiceLoss()(a.float(), b.float()), "metric": los.DiceMetric()(a.float(), b.float())}
    torch.save(out, OUT)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    sys.exit(main())
