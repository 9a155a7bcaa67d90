"""One dL/dvolume evaluation at config-2 geometry (512^3, 256^2, 8 poses) for ncu / timing:
    python scripts/prof_volgrad.py [trilinear|siddon] [poses]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import xvr_b200  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "trilinear"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
cfg = dict(bench.CONFIGS["trilinear"], renderer=kind)
drr = bench.build_scene(dev, cfg)
rot, xyz = bench.pose_batch(B, seed=0)
vol = drr.density.detach().clone().requires_grad_()
drr.density = vol
gout = torch.rand(B, 1, cfg["det"], cfg["det"], device=dev)
for it in range(2):
    vol.grad = None
    img = drr(xvr_b200.convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img.backward(gout)
    e1.record()
    torch.cuda.synchronize()
samples = B * cfg["det"] ** 2 * bench.N_POINTS
ms = e0.elapsed_time(e1)
print(f"{kind} dL/dvolume, {B} poses at {cfg['vol']}^3 / {cfg['det']}^2: {ms:.2f} ms "
      + (f"= {samples * 64 / ms / 1e6:.0f} GB/s of the 64 B/sample scatter model" if kind == "trilinear" else ""))
