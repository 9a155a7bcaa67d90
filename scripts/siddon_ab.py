"""A/B of the Siddon forward at config-5 geometry (768^3, 512^2): integer walk (default) vs certified evaluation of
every voxel index, with and without the Jacobian, fused entry; B = 32 poses per launch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import xvr_b200  # noqa: E402
from xvr_b200._lib import options  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(bench.CONFIGS["siddon"])
drr = bench.build_scene(dev, cfg)
B = int(os.environ.get("AB_BATCH", "32"))
rot, xyz = bench.pose_batch(B, seed=0)
rot, xyz = rot.to(dev), xyz.to(dev)
out = {}
for walk in (True, False):
    for grad in (False, True):
        with options(siddon_walk=walk):
            def run():
                r = rot.clone().requires_grad_(grad)
                x = xyz.clone().requires_grad_(grad)
                img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
                if grad:
                    img.sum().backward()
                return img
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                img = run()
            e1.record()
            torch.cuda.synchronize()
            out[f"{'walk' if walk else 'checked'}_{'fwd+bwd' if grad else 'fwd'}_ms_per_{B}"] = e0.elapsed_time(e1) / 5
            out[f"{'walk' if walk else 'checked'}_checksum"] = float(img.double().sum())
print(json.dumps(out))
