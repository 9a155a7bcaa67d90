#!/usr/bin/env python
"""Benchmark of the DRR hot path: DRRs/sec (trilinear forward + backward w.r.t. the pose), 512^3 CT, 256x256
detector, batch of 116 poses per GPU (BASELINE.json configs[1]), with the kernel's HBM roofline and the
reference's CPU path (the oracle restatement on real grid_sample) timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); poses shard across ranks with no data-path collective
(weak scaling: every rank renders its own 116 poses of a replicated volume).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

VOL_N = 512
DET = 256
BATCH = 116
N_POINTS = 500
SDD = 1020.0
DELX = 1.08821875
POSE_RANGES = dict(alphamin=-45, alphamax=45, betamin=-45, betamax=45, gammamin=-15, gammamax=15, txmin=-50,
                   txmax=50, tymin=700, tymax=900, tzmin=-50, tzmax=50)
METRIC = "DRRs/sec (fwd+bwd) 512^3 vol @256^2 det"
UNIT = "DRR/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, batch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), scaled
    to this batch; None when no capture of this kernel exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)[kernel]
        return rec["dram_bytes_per_launch"] * batch / rec["batch"]
    except Exception:
        return None


def algorithmic_bytes_fwd(h, w, n_points):
    """SURVEY.md 8(d) gather model: 8 corner voxels x 4 B per sample + the output pixel."""
    return h * w * (n_points * 32 + 4)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ scene
def pose_batch(batch, seed):
    from xvr_b200.sampler import random_pose_params

    g = torch.Generator().manual_seed(seed)
    rot, xyz = random_pose_params(**POSE_RANGES, batch_size=batch, generator=g)
    return torch.deg2rad(rot), xyz


def build_scene(device, vol_n=VOL_N, det=DET):
    import xvr_b200
    from xvr_b200.data import read, synthetic_ct

    hu, _, affine = synthetic_ct(vol_n, seed=0, device=device)
    sub = read(hu, affine=affine)
    del hu
    drr = xvr_b200.DRR(sub, SDD, det, DELX * 256.0 / det, renderer="trilinear", reverse_x_axis=False).to(device)
    return drr


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step(density, affinv, reorient, rot, xyz, rows, det, gout):
    """One fwd+bwd(pose) of the oracle (DiffDRR glue on real grid_sample) over a subset of detector rows."""
    import oracle

    rot = rot.clone().requires_grad_()
    xyz = xyz.clone().requires_grad_()
    pose = oracle.pose_from_params(rot, xyz, "euler_angles", "ZXY")
    src, tgt = oracle.detector_rays(pose, reorient, det, det, DELX * 256.0 / det, DELX * 256.0 / det, 0.0, 0.0, SDD, False)
    tgt = tgt.view(len(rot), det, det, 3)[:, rows].reshape(len(rot), -1, 3)
    raylen = (tgt - src).norm(dim=-1).unsqueeze(1)
    src, tgt = oracle.apply(affinv, src), oracle.apply(affinv, tgt)
    img = oracle.trilinear_render(density, src, tgt, raylen, n_points=N_POINTS)
    (img * gout).sum().backward()
    return img.detach(), rot.grad, xyz.grad


def cpu_reference(steps, warmup, vol_n=VOL_N, det=DET, min_seconds=0.0, max_seconds=120.0):
    """Time the reference's CPU path on this box's host cores on a bounded sample of the workload.

    One step = `b` poses x 8 of the detector's rows (fwd + bwd w.r.t. the pose through autograd).  `steps` timed
    steps are run after `warmup` untimed ones; the count is raised until `min_seconds` of CPU work have been timed
    (the in-line `cpu_baseline` of the main arm asks for ~12 s) and cut so that the timed region stays under
    `max_seconds` (the `--impl reference` arm must end within a few minutes whatever --steps says)."""
    import numpy as np

    from xvr_b200.data import REORIENT, read, synthetic_ct

    threads = torch.get_num_threads()
    hu, _, affine = synthetic_ct(vol_n, seed=0)
    sub = read(hu, affine=affine)
    density = sub.density
    affinv = torch.as_tensor(np.linalg.inv(sub.volume.affine), dtype=torch.float32)[None]
    reorient = torch.tensor(REORIENT["AP"])
    # ATen's CPU grid_sampler_3d parallelises over the batch dimension only -> one pose per thread
    b = max(4, min(threads, 32, BATCH))
    rot, xyz = pose_batch(BATCH, seed=0)
    rot, xyz = rot[:b], xyz[:b]
    n_rows = max(1, det // 32)
    rows = torch.arange(0, det, det // n_rows)[:n_rows]
    gout = torch.rand(b, 1, len(rows) * det, generator=torch.Generator().manual_seed(1))
    for _ in range(max(warmup, 1)):
        cpu_reference_step(density, affinv, reorient, rot, xyz, rows, det, gout)
    times = []
    while True:
        t0 = time.perf_counter()
        cpu_reference_step(density, affinv, reorient, rot, xyz, rows, det, gout)
        times.append(time.perf_counter() - t0)
        t = sum(times)
        per = t / len(times)
        if t + per > max_seconds or (len(times) >= steps and t >= min_seconds):
            break
    drr_equiv = b * len(rows) / det
    sample = (f"{len(times)} steps of {b} poses x {len(rows)} of {det} detector rows ({len(rows) * det} rays each, "
              f"{N_POINTS} samples/ray) of the {vol_n}^3 volume = {drr_equiv:.3f} DRR-equivalents per step, "
              f"{t:.1f} s of CPU work; fwd+bwd(pose) through autograd")
    return {"value": drr_equiv * len(times) / t, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "seconds_per_step": t / len(times), "steps": len(times)}


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--vol", type=int, default=VOL_N)
    ap.add_argument("--det", type=int, default=DET)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"{args.vol}^3 synthetic CT (fp32), batch={args.batch} poses per GPU, {args.det}x{args.det} detector, "
                f"trilinear n_points={N_POINTS}, fwd + bwd w.r.t. 6-DoF pose")
    vol_mib = args.vol ** 3 * 4 / 2 ** 20
    config = {"workload": workload, "renderer": "trilinear", "parallelism": f"pose-sharded x{world}",
              "l2_policy": (f"inputs larger than L2: the {vol_mib:.0f} MiB volume (plus its texture copy) is re-read "
                            "by every pose; no explicit flush" if vol_mib > 126 else
                            f"volume ({vol_mib:.0f} MiB) fits in L2: NOT a valid timing configuration"),
              "e2e_pipeline": "poses H2D from pinned memory, DRRs + pose gradients D2H to pinned memory every step; "
                              "double-buffered (copy of step i overlaps the render of step i+1, host reads step i-1)"}

    if args.impl == "reference":
        if rank != 0:
            return
        warmup = max(1, min(args.warmup, 2))
        res = cpu_reference(max(1, args.steps), warmup, args.vol, args.det, min_seconds=0.0, max_seconds=120.0)
        steps = res["steps"]
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    import xvr_b200
    from xvr_b200 import _lib

    drr = build_scene(device, args.vol, args.det)
    B, H, W = args.batch, args.det, args.det
    rot_h, xyz_h = pose_batch(B, seed=rank)
    rot_h, xyz_h = rot_h.pin_memory(), xyz_h.pin_memory()
    gout = torch.rand(B, 1, H, W, device=device, generator=torch.Generator(device=device).manual_seed(1))
    img_h = torch.empty(B, 1, H, W, pin_memory=True)
    grad_h = torch.empty(B, 6, pin_memory=True)

    def step(rot, xyz):
        rot = rot.detach().requires_grad_()
        xyz = xyz.detach().requires_grad_()
        pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
        img = drr(pose)
        img.backward(gout)
        return img, rot.grad, xyz.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rot_d, xyz_d = rot_h.to(device), xyz_h.to(device)
    for _ in range(max(args.warmup, 3)):
        step(rot_d, xyz_d)
    barrier()

    # ---- device-resident timing ("value")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.start_profile()
    launches0 = _lib.lib().xvr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed steps
    e0.record()
    for _ in range(args.steps):
        step(rot_d, xyz_d)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().xvr_launch_count() - launches0
    kernel_ms = _lib.stop_profile()

    # ---- end-to-end timing through the public API with host buffers ("e2e"): every step copies its poses in from
    # pinned memory and its DRRs + pose gradients out to pinned memory.  Double-buffered: the device->host copy of
    # step i runs on a copy stream while step i+1 renders, and the host consumes (waits for) the result of step
    # i-1 before it queues step i+1 -- every copy and every wait is inside the timed region.
    img_hh = [img_h, torch.empty_like(img_h).pin_memory()]
    grad_hh = [grad_h, torch.empty_like(grad_h).pin_memory()]
    copy_stream = torch.cuda.Stream(device=device)
    done = [torch.cuda.Event(), torch.cuda.Event()]
    main_stream = torch.cuda.current_stream()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    keep = [None, None]
    checksum = 0.0
    for i in range(args.steps):
        r, x = rot_h.to(device, non_blocking=True), xyz_h.to(device, non_blocking=True)
        img, gr, gx = step(r, x)
        g6 = torch.cat([gr, gx], dim=1)  # one contiguous (B,6) block: a strided D2H copy would be staged + synchronous
        ready = torch.cuda.Event()
        ready.record(main_stream)
        slot = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            img_hh[slot].copy_(img, non_blocking=True)
            grad_hh[slot].copy_(g6, non_blocking=True)
            done[slot].record(copy_stream)
        keep[slot] = (img, g6)  # keep the device tensors alive until their copy has run
        if i > 0:  # the caller reads the previous step's result while this one renders
            done[slot ^ 1].synchronize()
            checksum += float(img_hh[slot ^ 1][0, 0, H // 2, W // 2]) + float(grad_hh[slot ^ 1][0, 0])
    done[(args.steps - 1) & 1].synchronize()
    checksum += float(img_hh[(args.steps - 1) & 1][0, 0, H // 2, W // 2])
    e3.record()
    barrier()
    assert checksum == checksum, "e2e result is NaN"
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        total = world * B * args.steps
        value = total / (ms * 1e-3)
        peak, peak_src = peaks()
        name = next((k for k in ("xvr_trilinear_drr_fwd", "xvr_trilinear_drr_fwd_staged") if k in kernel_ms),
                    "xvr_trilinear_rays_fwd")
        k_ms = kernel_ms.get(name, [])
        k_avg = sum(k_ms) / len(k_ms) if k_ms else float("nan")
        alg = B * algorithmic_bytes_fwd(H, W, N_POINTS)
        achieved = alg / (k_avg * 1e-3) / 1e9
        step_share = {k: sum(v) / ms for k, v in kernel_ms.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 6 * 4,
                    "d2h_bytes_per_step": B * H * W * 4 + B * 6 * 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "trilinear_fwd_kernel<JAC=true> (one gather pass yields the DRR "
                         "and its per-ray pose Jacobian; the backward is a 28 B/ray epilogue)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "entry_point": name,
                         "traffic": measured_traffic(name, B) if (H, W, args.vol) == (DET, DET, VOL_N) else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "kernel_ms": k_avg},
            "roofline_step": {"note": "SURVEY 8(d) fwd+bwd(pose) figure (two gather passes, 2.0977 GB/DRR) over the "
                              "whole step time", "achieved": 2 * alg * args.steps / (ms * 1e-3) / 1e9 / 1.0,
                              "frac": 2 * alg * args.steps / (ms * 1e-3) / 1e9 / peak},
            "kernel_share_of_step": step_share,
        }
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference(1, 1, args.vol, args.det, min_seconds=12.0, max_seconds=30.0)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
