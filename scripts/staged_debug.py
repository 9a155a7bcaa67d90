"""Staged kernel vs texture kernel on one small scene: where and by how much do they differ?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xvr_b200  # noqa: E402
from tests._scene import make_drr, pose_params  # noqa: E402
from xvr_b200 import renderers  # noqa: E402

n, h, b = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
drr = make_drr(n, h)
rot, xyz = pose_params(b, seed=31)
imgs = []
for staged in ("0", "1"):
    os.environ["XVR_B200_STAGED"] = staged
    stats = torch.zeros(3, dtype=torch.int64, device="cuda")
    renderers._staged_stats["tensor"] = stats if staged == "1" else None
    from xvr_b200._lib import options
    with torch.no_grad(), options(ksplit=0):
        img = drr(xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY"))
    torch.cuda.synchronize()
    imgs.append(img)
    print("staged", staged, "stats", stats.tolist(), "sum", float(img.double().sum()))
d = (imgs[0] - imgs[1]).abs()
print("max abs diff", float(d.max()), "rel", float(d.max() / imgs[0].abs().max()), "n differing", int((d > 0).sum()), "of", d.numel())
bad = (d > 0).nonzero()
print("first differing (b,c,i,j):", bad[:12].tolist())
for bb in range(b):
    m = d[bb, 0] > 0
    print("pose", bb, "rows with diffs", m.any(1).nonzero().flatten().tolist()[:40], "cols", m.any(0).nonzero().flatten().tolist()[:40])
