"""torch.profiler table of one xvr training step (config 4 shapes) on one GPU: where the step time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench, xvr_b200
from xvr_b200.data import read, synthetic_ct
from xvr_b200.pose import RigidTransform, convert
from xvr_b200.preprocess import XrayTransforms
from xvr_b200.trainer import PoseRegressor, TrainStep

height = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
volumes, drr = [], None
for seed in range(2):
    hu, lab, affine = synthetic_ct(256, seed=seed, device=dev)
    sub = read(hu, lab, affine=affine, center_volume=False)
    if drr is None:
        drr = xvr_b200.DRR(sub, bench.SDD, height, bench.DELX * 256.0 / height, renderer="trilinear", reverse_x_axis=False).to(dev)
        drr.density = None
    aff = torch.as_tensor(affine, dtype=torch.float32, device=dev)
    center = aff[:3, :3] @ ((torch.tensor(hu.shape, device=dev) - 1) / 2) + aff[:3, 3]
    offset = convert(torch.zeros(1, 3, device=dev), center[None], parameterization="euler_angles", convention="ZXY")
    volumes.append((hu, None, RigidTransform(torch.linalg.inv(aff)), offset))
torch.manual_seed(0)
model = PoseRegressor("resnet18", "quaternion_adjugate", "ZXY", height=height, norm_layer="groupnorm").to(dev)
step = TrainStep(drr, model, volumes, bench.POSE_RANGES, XrayTransforms(height), bench.SDD, batch_size=116, n_grad_accum_itrs=4, n_warmup_itrs=8)
for i in range(6):
    step.step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6, 10):
        step.step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
