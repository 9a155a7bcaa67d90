"""BASELINE config 4: xvr train step on on-the-fly DRRs from 8 synthetic 256^3 volumes, batch 116 sharded over the
ranks of one box.  Launch with torchrun for N > 1.  Prints one JSON line from rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench, xvr_b200
from xvr_b200.data import read, synthetic_ct
from xvr_b200.pose import RigidTransform, convert
from xvr_b200.preprocess import XrayTransforms
from xvr_b200.trainer import PoseRegressor, TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--vol", type=int, default=256)
ap.add_argument("--n-vols", type=int, default=8)
ap.add_argument("--height", type=int, default=128)
ap.add_argument("--batch", type=int, default=116)
ap.add_argument("--labels", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay the iteration as CUDA graphs (TrainStep(use_cuda_graph=True))")
ap.add_argument("--log-every", type=int, default=4)
ap.add_argument("--channels-last", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

volumes, drr = [], None
for seed in range(args.n_vols):
    hu, lab, affine = synthetic_ct(args.vol, seed=seed, with_labels=args.labels, device=dev)
    sub = read(hu, lab, affine=affine, center_volume=False)  # per-subject affine + isocenter offset, as Trainer.load
    if drr is None:
        drr = xvr_b200.DRR(sub, bench.SDD, args.height, bench.DELX * 256.0 / args.height, renderer="trilinear",
                           reverse_x_axis=False).to(dev)
        drr.density = None
    aff = torch.as_tensor(affine, dtype=torch.float32, device=dev)
    center = aff[:3, :3] @ ((torch.tensor(hu.shape, device=dev) - 1) / 2) + aff[:3, 3]
    offset = convert(torch.zeros(1, 3, device=dev), center[None], parameterization="euler_angles", convention="ZXY")
    volumes.append((hu, lab.float() if lab is not None else None, RigidTransform(torch.linalg.inv(aff)), offset))

torch.manual_seed(0)
model = PoseRegressor("resnet18", "quaternion_adjugate", "ZXY", height=args.height, norm_layer="groupnorm",
                      channels_last=args.channels_last).to(dev)
step = TrainStep(drr, model, volumes, bench.POSE_RANGES, XrayTransforms(args.height), bench.SDD, batch_size=args.batch,
                 n_grad_accum_itrs=4, n_warmup_itrs=8, use_cuda_graph=args.graph,
                 log_every=args.log_every if args.graph else 1)
i0 = 0
for i0 in range(args.warmup):
    log = step.step(i0)
i0 = args.warmup
while args.graph and (len(step._graphs) < args.n_vols or step._opt_graph is None):  # capture every subject's graph
    log = step.step(i0)
    i0 += 1
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
for i in range(i0, i0 + args.steps):
    log = step.step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    print(json.dumps({"workload": f"xvr train step: {args.n_vols} x {args.vol}^3 volumes, batch {args.batch} sharded over {world} GPU(s), "
                      f"{args.height}^2 DRRs, resnet18+GroupNorm, 2 renders/step, labels={args.labels}, cuda_graph={args.graph}, channels_last={args.channels_last}",
                      "n_gpus": world, "ms_per_step": 1e3 * dt / args.steps, "steps_per_s": args.steps / dt,
                      "drrs_per_s": 2 * args.batch * args.steps / dt, "last_log": log}))
if world > 1:
    if args.graph:
        # captured graphs hold NCCL kernels of this communicator: tearing the communicator down under them hangs
        # (observed with NCCL 2.28.9) -- leave it to process exit
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)
    dist.destroy_process_group()
