"""Generate the golden vectors of tests/golden/ with the oracle on the CPU.

    python tests/golden/make_golden.py

The reference tree holds no fixtures for this path and diffdrr cannot be imported in the build container
(SURVEY.md 8c), so these vectors pin the ORACLE (oracle/*.py at the knob settings recorded in the file), not
DiffDRR itself: "parity unpinned".  They guard against silent drift of the oracle and give the GPU tests a
device-independent target.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import knobs  # noqa: E402


def scene(n=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    ax = torch.linspace(-1, 1, n)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    hu = torch.full((n, n, n), -1000.0)
    soft = (X / 0.9) ** 2 + (Y / 0.7) ** 2 + (Z / 0.9) ** 2 < 1
    hu = torch.where(soft, 40.0 + 20.0 * torch.randn(n, n, n, generator=g), hu)
    bone = ((X - 0.2) / 0.3) ** 2 + ((Y + 0.1) / 0.2) ** 2 + ((Z - 0.1) / 0.4) ** 2 < 1
    hu = torch.where(bone, torch.full_like(hu, 1100.0), hu)
    labels = soft.to(torch.uint8) + bone.to(torch.uint8)
    sp = 256.0 / n
    affine = np.diag([sp, sp, sp, 1.0])
    affine[:3, 3] = -sp * (n - 1) / 2
    return hu, labels, affine


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    hu, labels, affine = scene()
    density = oracle.hu_to_density(hu, 1.0)
    affinv = torch.as_tensor(np.linalg.inv(affine), dtype=torch.float32)[None]
    rot = torch.tensor([[0.10, -0.20, 0.05], [-0.55, 0.30, -0.12], [0.70, 0.65, 0.20]])
    xyz = torch.tensor([[10.0, 780.0, -20.0], [-35.0, 850.0, 15.0], [25.0, 720.0, 40.0]])
    det = dict(height=24, width=20, delx=9.0, dely=10.0, x0=3.0, y0=-4.0, sdd=1020.0, reverse_x_axis=True)
    out = {
        "knobs": {k: getattr(knobs, k) for k in dir(knobs) if k.isupper()},
        "hu": hu, "labels": labels, "affine": torch.as_tensor(affine), "density": density, "rot": rot, "xyz": xyz,
        "detector": det,
    }
    for renderer in ("trilinear", "siddon"):
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        pose = oracle.pose_from_params(r, x, "euler_angles", "ZXY")
        img = oracle.drr_forward(density, affinv, pose, reorient=oracle.REORIENT["AP"], renderer=renderer, **det)
        w = torch.linspace(0.5, 1.5, img.numel()).view_as(img)
        (img * w).sum().backward()
        out[renderer] = {"img": img.detach(), "grad_rot": r.grad, "grad_xyz": x.grad}
        pose = oracle.pose_from_params(rot, xyz, "euler_angles", "ZXY")
        out[renderer]["img_channels"] = oracle.drr_forward(
            density, affinv, pose, reorient=oracle.REORIENT["AP"], renderer=renderer, mask=labels, **det)
    # similarity metrics on the two renderers' images
    a = oracle.xray_transforms(out["trilinear"]["img"], 24, 20)
    b = oracle.xray_transforms(out["siddon"]["img"].roll(1, -1), 24, 20)
    out["metrics"] = {
        "x1": a, "x2": b,
        "ncc": oracle.ncc(a, b), "ncc9": oracle.ncc(a, b, 9),
        "mncc": oracle.multiscale_ncc(a, b, (None, 9), (0.5, 0.5)),
        "gncc11": oracle.gradient_ncc(a, b, 11, 0.0),
    }
    # pose parameterisations: every convert() flavour on fixed inputs
    gen = torch.Generator().manual_seed(1)
    poses = {}
    for name, k in oracle.N_ANGULAR_COMPONENTS.items():
        p = torch.randn(4, k, generator=gen) * 0.4
        if name in ("quaternion", "quaternion_adjugate", "rotation_10d"):
            p = p + torch.eye(k)[0]
        t = torch.randn(4, 3, generator=gen) * 50
        poses[name] = {"rot": p, "xyz": t,
                       "matrix": oracle.pose_from_params(p, t, name, "ZXY" if name == "euler_angles" else None)}
    out["poses"] = poses
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
