#!/usr/bin/env python
"""Benchmark of the DRR hot path (BASELINE.json).

Headline line = configs[1]: DRRs/sec, trilinear forward + backward w.r.t. the 6-DoF pose, 512^3 CT, 256x256 detector,
batch of 116 poses per GPU, with the dominant kernel's HBM roofline and the reference's CPU path (the oracle
restatement on real grid_sample) timed on the same box.  The same JSON line carries configs[4] under
"config5_siddon": Siddon renderer, 768^3 CT, 512x512 detector, batch of 256 poses per GPU, forward + backward(pose),
same keys (value, e2e, roofline, cpu_baseline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config both|trilinear|siddon]

N > 1 is launched by torchrun (one rank per GPU, NCCL); poses shard across ranks with no data-path collective
(weak scaling: every rank renders its own batch of a replicated volume).  `--config siddon` makes config 5 the
headline of the line instead (used for the per-config profiles under profiles/).
"""

import argparse
import json
import math
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

N_POINTS = 500
SDD = 1020.0
DELX = 1.08821875
POSE_RANGES = dict(alphamin=-45, alphamax=45, betamin=-45, betamax=45, gammamin=-15, gammamax=15, txmin=-50,
                   txmax=50, tymin=700, tymax=900, tzmin=-50, tzmax=50)
UNIT = "DRR/s"
CONFIGS = {
    # BASELINE.json configs[1] (the configuration `metric` is quoted on)
    "trilinear": dict(renderer="trilinear", vol=512, det=256, batch=116,
                      metric="DRRs/sec (fwd+bwd) 512^3 vol @256^2 det"),
    # BASELINE.json configs[4]
    "siddon": dict(renderer="siddon", vol=768, det=512, batch=256,
                   metric="DRRs/sec (fwd+bwd) Siddon 768^3 vol @512^2 det"),
}
DEFAULT_STEPS = {"trilinear": 130, "siddon": 8}  # >= 2 s of timed region each when no --steps is given


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel, batch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), scaled
    to this batch -- NOT measured in this run (ncu cannot run inside a timed bench); None without a capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)[kernel]
        return rec["dram_bytes_per_launch"] * batch / rec["batch"], rec.get("source", "profiles/traffic.json")
    except Exception:
        return None, None


def profiled_limiter(kernel):
    """What the committed ncu --set full capture of `kernel` says bounds it (unit utilisations of that launch, per cent
    of peak) -- context for `roofline`, NOT measured in this run; None without a capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel].get("limiter")
    except Exception:
        return None


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n)


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


def build_scene(device, cfg, det=None, renderer="trilinear"):
    """The DRR module of a configuration (`cfg` = an entry of CONFIGS, or a volume edge with `det` for the scripts)."""
    import xvr_b200
    from xvr_b200.data import read, synthetic_ct

    if not isinstance(cfg, dict):
        cfg = dict(vol=int(cfg), det=int(det), renderer=renderer)

    hu, _, affine = synthetic_ct(cfg["vol"], seed=0, device=device)
    sub = read(hu, affine=affine)
    del hu
    return xvr_b200.DRR(sub, SDD, cfg["det"], DELX * 256.0 / cfg["det"], renderer=cfg["renderer"],
                        reverse_x_axis=False).to(device)


def workload_text(cfg):
    how = f"trilinear n_points={N_POINTS}" if cfg["renderer"] == "trilinear" else "Siddon exact traversal"
    return (f"{cfg['vol']}^3 synthetic CT (fp32), batch={cfg['batch']} poses per GPU, {cfg['det']}x{cfg['det']} "
            f"detector, {how}, fwd + bwd w.r.t. 6-DoF pose")


def siddon_segment_count(drr, rot, xyz, chunk=8, trim=False):
    """Total number of traversed segments of the batch (the Siddon algorithmic-bytes figure of SURVEY 8(d)), counted
    by the traversal kernel itself in count-only mode on materialised rays, a few poses at a time."""
    import xvr_b200
    from xvr_b200._lib import call, opts_word, ptr, stream

    total = 0
    handle = drr.renderer._texture.get(drr.density) if trim else None  # occupancy: the trimmed traversal
    for i in range(0, rot.shape[0], chunk):
        pose = xvr_b200.convert(rot[i:i + chunk], xyz[i:i + chunk], parameterization="euler_angles", convention="ZXY")
        source, target = drr.detector(pose, None)
        source, target = drr.affine_inverse(source).contiguous(), drr.affine_inverse(target).contiguous()
        B, N, _ = target.shape
        cnt = torch.zeros(B, N, dtype=torch.int32, device=target.device)
        call("xvr_siddon_trace", ptr(drr.density), handle, *drr.density.shape, ptr(source), ptr(target), B, N,
             float(drr.renderer.voxel_shift), float(drr.renderer.eps), 0, None, None, ptr(cnt), opts_word(), stream())
        total += int(cnt.sum(dtype=torch.int64).item())
    return total


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step(cfg, density, affinv, reorient, rot, xyz, rows, gout):
    """One fwd+bwd(pose) of the oracle (DiffDRR glue on real grid_sample / sort) over a subset of detector rows."""
    import oracle

    det = cfg["det"]
    rot = rot.clone().requires_grad_()
    xyz = xyz.clone().requires_grad_()
    pose = oracle.pose_from_params(rot, xyz, "euler_angles", "ZXY")
    pix = DELX * 256.0 / det
    src, tgt = oracle.detector_rays(pose, reorient, det, det, pix, pix, 0.0, 0.0, SDD, False)
    tgt = tgt.view(len(rot), det, det, 3)[:, rows].reshape(len(rot), -1, 3)
    raylen = (tgt - src).norm(dim=-1).unsqueeze(1)
    src, tgt = oracle.apply(affinv, src), oracle.apply(affinv, tgt)
    if cfg["renderer"] == "trilinear":
        img = oracle.trilinear_render(density, src, tgt, raylen, n_points=N_POINTS)
    else:
        img = oracle.siddon_render(density, src, tgt, raylen)
    (img * gout).sum().backward()
    return img.detach(), rot.grad, xyz.grad


def cpu_reference(cfg, steps, warmup, min_seconds=0.0, max_seconds=150.0):
    """Time the reference's CPU path on this box's host cores on a bounded sample of the workload.

    One step = `b` poses x a few of the detector's rows (fwd + bwd w.r.t. the pose through autograd).  `warmup`
    untimed steps, then `steps` timed ones; the count is raised until `min_seconds` of CPU work have been timed (the
    in-line `cpu_baseline` of the main arm asks for ~12 s) and cut so that warm-up + timed region stay under
    `max_seconds` (the `--impl reference` arm must end within a few minutes whatever --steps says)."""
    import numpy as np

    from xvr_b200.data import REORIENT, read, synthetic_ct

    threads = host_threads()
    torch.set_num_threads(threads)
    det = cfg["det"]
    hu, _, affine = synthetic_ct(cfg["vol"], seed=0)
    sub = read(hu, affine=affine)
    density = sub.density
    affinv = torch.as_tensor(np.linalg.inv(sub.volume.affine), dtype=torch.float32)[None]
    reorient = torch.tensor(REORIENT["AP"])
    # ATen's CPU grid_sampler_3d parallelises over the batch dimension only -> one pose per thread
    b = max(4, min(threads, 32, cfg["batch"]))
    rot, xyz = pose_batch(cfg["batch"], seed=0)
    rot, xyz = rot[:b], xyz[:b]
    # trilinear: 8 of 256 rows; Siddon materialises 3(D+1) crossings per ray (2 307 at 768^3): 2 of 512 rows
    n_rows = max(1, det // 32) if cfg["renderer"] == "trilinear" else 2
    rows = torch.arange(0, det, det // n_rows)[:n_rows]
    gout = torch.rand(b, 1, len(rows) * det, generator=torch.Generator().manual_seed(1))
    t_start = time.perf_counter()
    done_warm = 0
    for _ in range(max(warmup, 1)):
        cpu_reference_step(cfg, density, affinv, reorient, rot, xyz, rows, gout)
        done_warm += 1
        if time.perf_counter() - t_start > 0.4 * max_seconds:
            break
    times = []
    while True:
        t0 = time.perf_counter()
        cpu_reference_step(cfg, density, affinv, reorient, rot, xyz, rows, gout)
        times.append(time.perf_counter() - t0)
        t = sum(times)
        per = t / len(times)
        if (time.perf_counter() - t_start) + per > max_seconds or (len(times) >= steps and t >= min_seconds):
            break
    drr_equiv = b * len(rows) / det
    what = f"{N_POINTS} samples/ray" if cfg["renderer"] == "trilinear" else "sort-based Siddon"
    sample = (f"{len(times)} steps of {b} poses x {len(rows)} of {det} detector rows ({len(rows) * det} rays each, "
              f"{what}) of the {cfg['vol']}^3 volume = {drr_equiv:.4f} DRR-equivalents per step, "
              f"{t:.1f} s of CPU work on {threads} threads; fwd+bwd(pose) through autograd")
    return {"value": drr_equiv * len(times) / t, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "seconds_per_step": t / len(times), "steps": len(times), "warmup": done_warm}


# ------------------------------------------------------------------------------------------ GPU arm
def run_config(name, steps, warmup, rank, world, device, with_clocks=True):
    """Device-resident and end-to-end timing of one configuration; returns the dict of the JSON line (rank 0 fills
    the rank-independent parts after the max-over-ranks reduction)."""
    import torch.distributed as dist

    import xvr_b200
    from xvr_b200 import _lib

    cfg = CONFIGS[name]
    drr = build_scene(device, cfg)
    B, H, W = cfg["batch"], cfg["det"], cfg["det"]
    rot_h, xyz_h = pose_batch(B, seed=rank)
    rot_h, xyz_h = rot_h.pin_memory(), xyz_h.pin_memory()
    gout = torch.rand(B, 1, H, W, device=device, generator=torch.Generator(device=device).manual_seed(1))

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
    for _ in range(warmup):
        step(rot_d, xyz_d)
    barrier()

    # ---- empty-space trimming (trilinear): the kernel skips samples outside the box of the volume's non-zero voxels
    # (exact zeros, bit-identical output).  Measure what share of the samples it marches, and time the step once more
    # with the trimming switched off so that both numbers are on the line.
    trimming = None
    if name == "trilinear":
        import ctypes

        from xvr_b200._lib import call, options, opts_word, ptr, stream
        pose = xvr_b200.convert(rot_d, xyz_d, parameterization="euler_angles", convention="ZXY")
        cam2world = drr.detector.reorient.compose(pose).matrix
        cam2vox = (drr._affine_inverse @ cam2world)[:, :3].contiguous()
        cam2world = cam2world[:, :3].contiguous()
        origin, row_step, col_step = drr.detector.pixel_basis()
        det9 = (ctypes.c_float * 9)(*origin, *row_step, *col_step)
        handle = drr.renderer._texture.get(drr.density)
        counts = []
        for trim in (True, False):
            counter = torch.zeros(1, dtype=torch.int64, device=device)
            with options(trim=trim):
                call("xvr_trilinear_drr_count", handle, *drr.density.shape, ptr(cam2vox), ptr(cam2world), det9, B, H, W,
                     N_POINTS, float(drr.renderer.eps), ptr(counter), opts_word(), stream())
            counts.append(int(counter.item()))
    else:
        import ctypes

        from xvr_b200._lib import call, options, stream
        handle = drr.renderer._texture.get(drr.density)
        counts = [siddon_segment_count(drr, rot_d, xyz_d, trim=True), siddon_segment_count(drr, rot_d, xyz_d)]
    if True:
        bbox = (ctypes.c_int * 6)()
        call("xvr_volume_bbox", handle, bbox, stream())
        k = max(3, min(steps, 10))
        with options(trim=False):
            for _ in range(2):
                step(rot_d, xyz_d)
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(k):
                step(rot_d, xyz_d)
            t1.record()
            barrier()
        t_off = torch.tensor([t0.elapsed_time(t1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_off, op=dist.ReduceOp.MAX)
        unit = "samples" if name == "trilinear" else "segments"
        what = ("samples whose 8 corners lie outside the box of the volume's non-zero voxels / before the first and "
                "after the last occupied 8^3 brick on the ray (brick distance field, sphere-traced from both ends) are skipped (exact zeros for every sum: images and "
                "Jacobians bit-identical to the full march, "
                "tests/test_trilinear_gpu.py::test_empty_space_trimming_is_bit_identical)" if name == "trilinear" else
                "plane crossings before a ray enters the first occupied 8^3 brick and after it leaves the last one are "
                "dropped (the segments are air: exact zeros for the line integral and the Jacobian sums, images and "
                "Jacobians bit-identical to the full traversal, "
                "tests/test_siddon_gpu.py::test_empty_space_trimming_is_bit_identical, "
                "::test_trimmed_traversal_is_a_run_of_the_full_one_and_drops_only_air)")
        trimming = {"what": what,
                    "nonzero_box": list(bbox), f"{unit}_in_volume": counts[1], f"{unit}_marched": counts[0],
                    "marched_fraction": counts[0] / max(1, counts[1]),
                    "value_without_trimming": world * B * k / (t_off.item() * 1e-3),
                    "ms_per_step_without_trimming": t_off.item() / k}

    # ---- device-resident timing ("value")
    sampler = ClockSampler(device.index)
    if rank == 0 and with_clocks:
        sampler.start()
    _lib.start_profile()
    launches0 = _lib.lib().xvr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed steps
    e0.record()
    for _ in range(steps):
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
    img_hh = [torch.empty(B, 1, H, W, pin_memory=True) for _ in range(2)]
    grad_hh = [torch.empty(B, 6, pin_memory=True) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)
    done = [torch.cuda.Event(), torch.cuda.Event()]
    main_stream = torch.cuda.current_stream()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    keep = [None, None]
    checksum = 0.0
    for i in range(steps):
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
            checksum += float(img_hh[slot ^ 1][0, 0, H // 2, W // 2].detach()) + float(grad_hh[slot ^ 1][0, 0])
    done[(steps - 1) & 1].synchronize()
    checksum += float(img_hh[(steps - 1) & 1][0, 0, H // 2, W // 2].detach())
    e3.record()
    barrier()
    assert checksum == checksum, "e2e result is NaN"
    ms_e2e = e2.elapsed_time(e3)

    # ---- sustained: the same device-resident step for >= 2 s.  The K-step region above lasts a fraction of a second --
    # too short to show that clocks and power have settled -- so the line also carries this longer run (not the
    # headline: `value` is exactly K steps, as the contract says); the clock sampler covers it.
    n_sus = torch.tensor([min(3000, max(steps, math.ceil(2000.0 * steps / max(ms, 1e-3))))], device=device)
    if world > 1:
        dist.all_reduce(n_sus, op=dist.ReduceOp.MAX)
    n_sus = int(n_sus.item())
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e4.record()
    for _ in range(n_sus):
        step(rot_d, xyz_d)
    e5.record()
    barrier()
    ms_sus = e4.elapsed_time(e5)
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None

    # every rank rendered its own poses: a non-finite or all-zero image anywhere invalidates the run
    ok = torch.tensor([float(torch.isfinite(img).all() and img.abs().sum() > 0)], device=device)
    t = torch.tensor([ms, ms_e2e, ms_sus], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    assert ok.item() == 1.0, "a rank rendered a non-finite or empty image batch"
    ms, ms_e2e, ms_sus = t.tolist()
    if rank != 0:
        return None

    total = world * B * steps
    peak, peak_src = peaks()
    if name == "trilinear":
        entry = next((k for k in ("xvr_trilinear_drr_fwd", "xvr_trilinear_drr_fwd_staged") if k in kernel_ms),
                     "xvr_trilinear_rays_fwd")
        # SURVEY 8(d) gather model: 8 corners x 4 B per sample + the pixel -- over the samples the kernel MARCHES (the
        # trimmed ones are never fetched; counting them would credit the kernel with bandwidth it does not use)
        marched = trimming["samples_marched"] if trimming else B * H * W * N_POINTS
        alg = marched * 32 + B * H * W * 4
        alg_note = (f"samples_marched*32 + B*H*W*4: 8 corner voxels x 4 B per MARCHED sample ({marched} of "
                    f"{B * H * W * N_POINTS}: empty-space trimming) + the output pixels (gather model, no reuse)")
        kernel = ("trilinear forward (+ per-ray pose Jacobian in the same march; the backward is a 28 B/ray epilogue)")
        two_pass = 2.0
    else:
        entry = "xvr_siddon_drr_fwd"
        n_seg = trimming["segments_marched"]  # the trimmed ones are never fetched: not credited
        alg = n_seg * 4 + B * H * W * 4  # one 4-byte voxel per traversed segment + the output pixel
        alg_note = (f"sum over rays of (n_seg*4+4): {n_seg} WALKED segments in the batch ({n_seg / (B * H * W):.1f} per "
                    f"ray; {trimming['segments_in_volume']} before the empty-space trimming), counted by "
                    "xvr_siddon_trace with the same occupancy handle")
        kernel = "siddon_fwd_kernel<JAC=true> (traversal + per-ray pose Jacobian in one pass)"
        two_pass = 2.0
    k_ms = kernel_ms.get(entry, [])
    k_avg = sum(k_ms) / len(k_ms) if k_ms else float("nan")
    achieved = alg / (k_avg * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(entry, B)
    vol_mib = cfg["vol"] ** 3 * 4 / 2 ** 20
    return {
        "metric": cfg["metric"], "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(cfg), "renderer": cfg["renderer"],
                   "parallelism": f"pose-sharded x{world}",
                   "l2_policy": f"inputs larger than L2: the {vol_mib:.0f} MiB volume is re-read by every pose; "
                                "no explicit flush",
                   "e2e_pipeline": "poses H2D from pinned memory, DRRs + pose gradients D2H to pinned memory every "
                                   "step; double-buffered (copy of step i overlaps the render of step i+1, host "
                                   "reads step i-1)"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 6 * 4,
                "d2h_bytes_per_step": B * H * W * 4 + B * 6 * 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "timed_seconds": ms * 1e-3,
        "sustained": {"steps": n_sus, "seconds": ms_sus * 1e-3, "value": world * B * n_sus / (ms_sus * 1e-3),
                      "unit": UNIT, "note": "the same device-resident step repeated for >= 2 s (clocks / power are "
                                            "sampled over the timed, e2e and sustained regions)"},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "entry_point": entry, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                     "algorithmic_bytes": alg_note, "kernel_ms": k_avg,
                     "limiter_from_profile": profiled_limiter(entry)},
        "roofline_step": {"note": "SURVEY 8(d) fwd+bwd(pose) figure (two gather passes) over the whole step time",
                          "achieved": two_pass * alg * steps / (ms * 1e-3) / 1e9,
                          "frac": two_pass * alg * steps / (ms * 1e-3) / 1e9 / peak},
        "kernel_share_of_step": {k: sum(v) / ms for k, v in kernel_ms.items()},
        **({"empty_space_trimming": trimming} if trimming else {}),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="both", choices=["both", "trilinear", "siddon"])
    ap.add_argument("--vol", type=int, default=None, help="override the volume edge (debugging; not a bench value)")
    ap.add_argument("--det", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    names = ["trilinear", "siddon"] if args.config == "both" else [args.config]
    for n in names:
        for k in ("vol", "det", "batch"):
            if getattr(args, k) is not None:
                CONFIGS[n][k] = getattr(args, k)
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (the oracle port: DiffDRR itself is not installable
        # here), all host threads, rank 0 only
        if rank != 0:
            return
        head = names[0]
        cfg = CONFIGS[head]
        steps = args.steps if args.steps is not None else 10
        res = cpu_reference(cfg, max(1, steps), args.warmup, min_seconds=0.0, max_seconds=150.0)
        line = {"impl": "reference", "metric": cfg["metric"], "value": res["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": res["steps"], "warmup": res["warmup"],
                "ms_per_step": res["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_text(cfg), "renderer": cfg["renderer"],
                           "parallelism": f"{res['cores']} host threads, one pose per thread"},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
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

    results = {}
    for i, n in enumerate(names):
        steps = args.steps if args.steps is not None else DEFAULT_STEPS[n]
        if i > 0 and args.steps is not None:
            # the secondary configuration is ~20x the work per step: a share of K that keeps the run within minutes
            steps = max(3, min(args.steps, 10))
        results[n] = run_config(n, steps, warmup, rank, world, device)
        torch.cuda.empty_cache()

    if rank == 0:
        line = results[names[0]]
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference(CONFIGS[names[0]], 1, 1, min_seconds=12.0, max_seconds=40.0)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        for n in names[1:]:
            sub = results[n]
            if world == 1 and not args.no_cpu_baseline:
                res = cpu_reference(CONFIGS[n], 1, 1, min_seconds=10.0, max_seconds=40.0)
                sub["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["config5_siddon" if n == "siddon" else f"config_{n}"] = sub
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
