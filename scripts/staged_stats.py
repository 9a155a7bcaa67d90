"""Share of the samples the staged kernel serves from shared memory at config 2 (one launch of 116 poses)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["XVR_B200_STAGED"] = "1"
import bench  # noqa: E402
import xvr_b200  # noqa: E402
from xvr_b200 import renderers  # noqa: E402

dev = torch.device("cuda:0")
cfg = bench.CONFIGS["trilinear"]
drr = bench.build_scene(dev, cfg)
rot, xyz = bench.pose_batch(cfg["batch"], seed=0)
stats = torch.zeros(8, dtype=torch.int64, device=dev)
renderers._staged_stats["tensor"] = stats
with torch.no_grad():
    img = drr(xvr_b200.convert(rot.to(dev), xyz.to(dev), parameterization="euler_angles", convention="ZXY"))
torch.cuda.synchronize()
s, g, t, _, irregular, unstaged, ahead, miss = stats.tolist()
print(json.dumps({"from_shared": s, "from_global": g, "timeouts": t, "shared_fraction": s / max(1, s + g), "global_irregular_rays": irregular,
                  "global_unstaged_stage": unstaged, "global_ahead_of_ring": ahead, "global_box_miss": miss}))
