"""Time the label-channel render (trainer.py:288 with mask=seg, channels collapsed as at trainer.py:294) at config-2 size:
    [XVR_B200_LIB=variant.so] python scripts/time_labels.py [batch]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import xvr_b200  # noqa: E402
from tests._scene import make_drr, pose_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 116
drr = make_drr(512, 256, with_labels=True)
rot, xyz = pose_params(B, seed=0)
gout = torch.rand(B, 1, 256, 256, device="cuda")


def step(channels):
    r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    img = drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"), mask_to_channels=channels)
    img = img.sum(dim=1, keepdim=True)
    img.backward(gout)
    return img


res = []
for channels in (True, False):
    for _ in range(2):
        img = step(channels)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        step(channels)
    e1.record()
    torch.cuda.synchronize()
    res.append(f"{'label channels' if channels else 'single channel'}: {e0.elapsed_time(e1) / 5:.3f} ms "
               f"(sum {img.sum().item():.6e})")
print(os.path.basename(os.environ.get("XVR_B200_LIB", "default")), f"B={B}", " | ".join(res))
