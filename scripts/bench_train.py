"""BASELINE config 4: the xvr training step on on-the-fly DRRs from 8 synthetic 256^3 volumes, batch 116 sharded over the
ranks of one box (STRONG scaling: the global batch is fixed).  Launch with torchrun for N > 1.  Rank 0 prints one JSON
line; --profile adds the per-kernel GPU-time table of four steps.

The regressor's heads start at the MEAN of the pose distribution (--init near-truth, default): an untrained network
predicts a camera at the isocentre, its DRRs miss the volume and the renderer's early-out makes the second render of
every step free -- not the workload a training run spends its time on (round-1 finding)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench, xvr_b200
from xvr_b200.data import read, synthetic_ct
from xvr_b200.pose import RigidTransform, convert
from xvr_b200.preprocess import XrayTransforms
from xvr_b200.trainer import PoseRegressor, TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--vol", type=int, default=256)
ap.add_argument("--n-vols", type=int, default=8)
ap.add_argument("--height", type=int, default=128)
ap.add_argument("--batch", type=int, default=116)
ap.add_argument("--labels", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay the iteration as CUDA graphs (TrainStep(use_cuda_graph=True))")
ap.add_argument("--log-every", type=int, default=4)
ap.add_argument("--channels-last", action="store_true")
ap.add_argument("--bf16", action="store_true", help="bf16 autocast around the CNN (tensor-core convolutions)")
ap.add_argument("--init", default="near-truth", choices=["near-truth", "random"])
ap.add_argument("--profile", action="store_true")
ap.add_argument("--shard", default="batch", choices=["batch", "accumulation"],
                help="batch: every iteration split over all ranks; accumulation: the 4 iterations of an optimiser step "
                     "dealt out to groups of ranks (TrainStep docstring)")
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
                      channels_last=args.channels_last, autocast_bf16=args.bf16).to(dev)
if args.init == "near-truth":
    with torch.no_grad():
        mid = {k[:-3]: 0.5 * (bench.POSE_RANGES[k] + bench.POSE_RANGES[k[:-3] + "max"]) for k in bench.POSE_RANGES if k.endswith("min")}
        rot = torch.deg2rad(torch.tensor([[mid["alpha"], mid["beta"], mid["gamma"]]], device=dev))
        xyz = torch.tensor([[mid["tx"], mid["ty"], mid["tz"]]], device=dev)
        mean_pose = convert(rot, xyz, parameterization="euler_angles", convention="ZXY").compose(volumes[0][3])
        r10, t3 = mean_pose.convert("quaternion_adjugate")
        model.rot_regression.bias.copy_(r10[0])
        model.xyz_regression.bias.copy_(t3[0] / model.unit_conversion_factor)
        model.rot_regression.weight.mul_(0.05)
        model.xyz_regression.weight.mul_(0.05)
step = TrainStep(drr, model, volumes, bench.POSE_RANGES, XrayTransforms(args.height), bench.SDD, batch_size=args.batch,
                 n_grad_accum_itrs=4, n_warmup_itrs=8, use_cuda_graph=args.graph,
                 log_every=args.log_every if args.graph else 1, shard=args.shard)
i0 = 0
for i0 in range(args.warmup):
    log = step.step(i0)
i0 = args.warmup
def everything_captured():
    """Every rank has captured every subject's graph and the optimiser graph (ranks own different iterations when the
    accumulation window is sharded: agree on it collectively, at a window boundary)."""
    ok = torch.tensor([float(len(step._graphs) >= args.n_vols and step._opt_graph is not None)], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())


i0 += (-i0) % 4
while args.graph and not everything_captured():
    for _ in range(4):
        log = step.step(i0)
        i0 += 1
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for i in range(i0, i0 + args.steps):
    log = step.step(i)
e1.record()
torch.cuda.synchronize()
dt_dev = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(dt_dev, op=dist.ReduceOp.MAX)  # device time, max over ranks
    dist.barrier()
dt = time.time() - t0
table = None
if args.profile and rank == 0:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(i0 + args.steps, i0 + args.steps + 4):
            step.step(i)
        torch.cuda.synchronize()
    table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=70)
elif args.profile:
    for i in range(i0 + args.steps, i0 + args.steps + 4):
        step.step(i)
    torch.cuda.synchronize()
if rank == 0:
    ms = 1e3 * dt_dev.item() / args.steps
    print(json.dumps({"workload": f"xvr train step: {args.n_vols} x {args.vol}^3 volumes, global batch {args.batch} sharded over {world} GPU(s), "
                      f"{args.height}^2 DRRs, resnet18+GroupNorm, 2 renders/step, labels={args.labels}, cuda_graph={args.graph}, shard={args.shard}, "
                      f"channels_last={args.channels_last}, bf16={args.bf16}, init={args.init}",
                      "n_gpus": world, "ms_per_step": ms, "ms_per_step_wall": 1e3 * dt / args.steps, "steps_per_s": 1e3 / ms,
                      "drrs_per_s": 2 * args.batch * 1e3 / ms, "last_log": log}))
    if table:
        print(table)
if world > 1:
    if args.graph:
        # captured graphs hold NCCL kernels of this communicator: tearing the communicator down under them hangs
        # (observed with NCCL 2.28.9) -- leave it to process exit
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)
    dist.destroy_process_group()
