"""Time the C2 step (trilinear fwd + bwd(pose), B=116, 512^3, 256^2) or the C5 step (siddon, B given) for a list of warp /
CTA tile shapes (XVR_B200_TILE = lane_w_log2,cta_w_log2) with the library selected by XVR_B200_LIB.
    python scripts/sweep_tiles.py trilinear 0,3 1,3 2,3      python scripts/sweep_tiles.py siddon:32 0,3 2,3"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import xvr_b200  # noqa: E402

name, _, b = sys.argv[1].partition(":")
cfg = dict(bench.CONFIGS[name])
if b:
    cfg["batch"] = int(b)
dev = torch.device("cuda")
drr = bench.build_scene(dev, cfg)
rot, xyz = (t.to(dev) for t in bench.pose_batch(cfg["batch"], 0))
gout = torch.rand(cfg["batch"], 1, cfg["det"], cfg["det"], device=dev)


def step():
    r, x = rot.detach().requires_grad_(), xyz.detach().requires_grad_()
    img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
    img.backward(gout)
    return img


res = []
ref = None
from xvr_b200._lib import options  # noqa: E402

for spec in sys.argv[2:]:  # lane_w_log2,cta_w_log2[@ksplit]
    tile, _, ks = spec.partition("@")
    os.environ["XVR_B200_TILE"] = tile
    ctx = options(ksplit=int(ks)) if ks else options()
    ctx.__enter__()
    for _ in range(2):
        img = step()
    if ref is None:
        ref = img.detach().clone()
    same = bool(torch.equal(ref, img.detach()))
    n = 10 if name == "trilinear" else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ctx.__exit__(None, None, None)
    res.append(f"{spec}: {e0.elapsed_time(e1) / n:.3f} ms{'' if same else ' (rounding differs)'}")
print(os.path.basename(os.environ.get("XVR_B200_LIB", "default")), name, " | ".join(res))
