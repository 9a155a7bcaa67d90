"""Test-time pose optimisation: the host side of xvr's registration hot loop.

Mirrors ``_RegistrarBase.run`` / ``run_test_time_optimization`` of /root/reference/src/xvr/registrar/base.py:125-292
for inputs that are already tensors (reading DICOMs and predicting the initial pose with a CNN are outside the
hot path): multiscale Adam(maximize) on the ``Registration`` parameters with ``ReduceLROnPlateau(mode="max",
factor=0.1)``, similarity ``beta * mNCC([None, p]) + (1 - beta) * GradNCC(q, sigma)`` on ``XrayTransforms``-ed
images, a stage ending after ``max_n_plateaus`` learning-rate drops.

At B = 1 an iteration is ~0.2 ms of DRAM/TEX time, so launch overhead dominates the reference loop (~10^2 small
kernels and 3+ host syncs per iteration).  ``use_cuda_graph=True`` captures one whole iteration -- pose
conversion, fused DRR forward (+ Jacobian), transforms, similarity forward/backward, Adam update -- into a CUDA
graph and replays it; the only per-iteration host round trip left is the scalar the plateau scheduler needs.
"""

import ctypes
import time

import torch

from ._lib import call, ptr, stream
from .metrics import (GradientNormalizedCrossCorrelation2d, MultiscaleNormalizedCrossCorrelation2d,
                      RegistrationSimilarity)
from .pose import convert
from .preprocess import XrayTransforms
from .registration import Registration

__all__ = ["Registrar", "PlateauScheduler", "parse_scales", "adam_maximize_"]


def parse_scales(scales, crop, height):
    """Down-sampling factors -> per-stage rescale factors (base.py:402-407); ``scales`` is "8" or "8,4,2"."""
    if isinstance(scales, str):
        scales = scales.split(",")
    pyramid = [1.0] + [float(s) * (height / (height + crop)) for s in scales]
    return [pyramid[i] / pyramid[i + 1] for i in range(len(pyramid) - 1)]


class PlateauScheduler:
    """``torch.optim.lr_scheduler.ReduceLROnPlateau(mode="max", threshold_mode="rel", factor=...)`` plus the
    reference loop's stopping rule (base.py:262-278), evaluated ON THE DEVICE so that an iteration needs no host
    round trip: every quantity is a 0-dim tensor updated with ``torch.where``.

    ``step(metric)`` must be called after the optimiser update of the iteration, as the reference does.
    """

    def __init__(self, lrs, factor=0.1, patience=10, threshold=1e-4, min_lr=0.0, eps=1e-8, max_n_plateaus=3,
                 storage=None):
        dev = lrs[0].device
        self.lrs = lrs
        self.factor, self.patience, self.threshold, self.min_lr, self.eps = factor, patience, threshold, min_lr, eps
        self.max_n_plateaus = max_n_plateaus
        f64 = dict(device=dev, dtype=torch.float64)  # the reference's scheduler works in Python doubles
        if storage is None:
            storage = torch.zeros(5, **f64)
        # `storage` (5 doubles: best, num_bad, smallest lr seen, n_plateaus, active) lets the fused update kernel
        # (csrc/regstep.cu, xvr_reg_update) own the same state this class steps with tensor ops
        self.best, self.num_bad, self.current_lr, self.n_plateaus, self.active = (storage[i] for i in range(5))
        with torch.no_grad():
            self.best.fill_(float("-inf"))
            self.num_bad.zero_()
            self.current_lr.fill_(float("inf"))
            self.n_plateaus.zero_()
            self.active.fill_(1.0)  # 1 while the stage is running, 0 after the stopping rule fired

    def state(self):
        return [self.best, self.num_bad, self.current_lr, self.n_plateaus, self.active, *self.lrs]

    def step(self, metric):
        on = self.active > 0
        metric = metric.to(torch.float64)
        better = metric > self.best * (1.0 + self.threshold)
        best = torch.where(better, metric, self.best)
        num_bad = torch.where(better, torch.zeros_like(self.num_bad), self.num_bad + 1)
        reduce = num_bad > self.patience
        new_lrs = []
        for lr in self.lrs:
            cand = torch.clamp(lr * self.factor, min=self.min_lr)
            new_lrs.append(torch.where(reduce & (lr - cand > self.eps), cand, lr))
        num_bad = torch.where(reduce, torch.zeros_like(num_bad), num_bad)
        # reference bookkeeping: every time lr[0] falls below the smallest value seen, count a plateau
        dropped = new_lrs[0] < self.current_lr
        current = torch.where(dropped, new_lrs[0], self.current_lr)
        n_plateaus = self.n_plateaus + dropped.to(self.n_plateaus.dtype)
        active = torch.where(n_plateaus >= self.max_n_plateaus, torch.zeros_like(self.active), self.active)
        for dst, src in zip(self.state(), [best, num_bad, current, n_plateaus, active, *new_lrs]):
            dst.copy_(torch.where(on, src, dst))


def adam_maximize_(params, grads, state, lrs, active=None, betas=(0.9, 0.999), eps=1e-8):
    """One ``torch.optim.Adam(maximize=True)`` step written on tensors only (step count and learning rates live on
    the device) so that it can be captured in a CUDA graph.  ``active`` (0/1 tensor) masks the whole update."""
    b1, b2 = betas
    gate = 1.0 if active is None else active
    t = state["step"] + gate
    bias1 = 1 - b1**t
    bias2 = 1 - b2**t
    for p, g, lr, m, v in zip(params, grads, lrs, state["exp_avg"], state["exp_avg_sq"]):
        g = -g
        m_new = torch.lerp(m, g, 1 - b1)
        v_new = v * b2 + (1 - b2) * g * g
        denom = v_new.sqrt() / bias2.sqrt().to(p.dtype) + eps
        p_new = p - (lr / bias1).to(p.dtype) * (m_new / denom)
        if active is None:
            m.copy_(m_new), v.copy_(v_new), p.copy_(p_new)
        else:
            on = active > 0
            m.copy_(torch.where(on, m_new, m)), v.copy_(torch.where(on, v_new, v)), p.copy_(torch.where(on, p_new, p))
    state["step"].copy_(t)


class Registrar:
    """Multiscale intensity-based 2D/3D registration of one X-ray to a DRR module.

    ``run(gt, init_pose, intrinsics)`` returns ``(final_pose, info)``; ``info`` holds ``params`` (ZXY Euler angles +
    translation per iteration), ``nccs``, ``alphas`` (learning rates) and ``times`` -- the four lists of the
    reference's ``run_test_time_optimization`` -- plus ``runtime`` and ``n_itrs`` per stage.  Per-iteration wall
    times are the mean of the polling chunk they belong to (there is no per-iteration host synchronisation).
    """

    def __init__(self, drr, scales="8", n_itrs="500", parameterization="euler_angles", convention="ZXY", lr_rot=1e-2,
                 lr_xyz=1e0, patience=10, threshold=1e-4, max_n_plateaus=3, crop=0, equalize=False, mncc_patch_size=9,
                 gncc_patch_size=11, sigma=0.0, beta=0.5, use_cuda_graph=True, poll_every=16, fused_update=True,
                 fused_similarity=True, provenance=None, saveimg=False):
        self.drr = drr
        # what the reference's registrars know from their constructor arguments and `save` writes into
        # parameters.pt (base.py:355-394): volume / mask paths, labels, orientation, reverse_x_axis, renderer, ...
        self.provenance = dict(provenance or {})
        self.saveimg = saveimg
        self.scales = scales.split(",") if isinstance(scales, str) else [str(s) for s in scales]
        self.n_itrs = [int(n) for n in n_itrs.split(",")] if isinstance(n_itrs, str) else [int(n) for n in n_itrs]
        if len(self.scales) != len(self.n_itrs):
            raise ValueError("scales and n_itrs must have the same number of stages")
        self.parameterization, self.convention = parameterization, convention
        self.lr_rot, self.lr_xyz = lr_rot, lr_xyz
        self.patience, self.threshold, self.max_n_plateaus = patience, threshold, max_n_plateaus
        self.crop, self.equalize = crop, equalize
        self.beta = beta
        self.patches, self.sigma = (mncc_patch_size, gncc_patch_size), sigma
        # XrayTransforms + similarity + their backward as nine launches instead of ~60 (csrc/ncc.cu, xvr_regsim; the
        # ends that touch every pixel run as thread-block clusters): 0.378 vs 0.473 ms per iteration at config 3 on the
        # B200.  Covers the reference's defaults (no Equalize, sigma = 0); anything else takes the unfused chain.
        self.fused_similarity = bool(fused_similarity) and not equalize and sigma == 0.0
        self.sim1 = MultiscaleNormalizedCrossCorrelation2d([None, mncc_patch_size], [0.5, 0.5])
        self.sim2 = GradientNormalizedCrossCorrelation2d(gncc_patch_size, sigma)
        self.use_cuda_graph = use_cuda_graph
        self.poll_every = max(1, int(poll_every))
        # one launch for Adam + plateau scheduler + stopping rule + trajectory row instead of ~10^2 tensor ops
        self.fused_update = fused_update

    def imagesim(self, x, y):
        return self.beta * self.sim1(x, y) + (1 - self.beta) * self.sim2(x, y)

    def _score(self, transform, img, raw):
        """Scalar similarity of the raw DRR batch ``raw`` to the transformed target ``img`` (summed over the batch)."""
        if self.fused_similarity:
            if getattr(self, "_fsim_for", None) is not img:  # one fixed image (and its Sobel) per stage
                self._fsim = RegistrationSimilarity(img, *self.patches, beta=self.beta, mean=transform.mean,
                                                    std=transform.std, std_eps=transform.standardize.eps,
                                                    eps=self.sim1.eps)
                self._fsim_for = img
            return self._fsim(raw)
        return self.imagesim(img, transform(raw)).sum()

    # ------------------------------------------------------------------ one iteration (device only)
    def _iteration(self, reg, transform, img, state, sched, log):
        """forward -> similarity -> backward -> Adam(maximize) -> plateau logic -> trajectory row; no host sync."""
        reg.rotation.grad = None
        reg.translation.grad = None
        loss = self._score(transform, img, reg())
        loss.backward()
        if self.fused_update:
            hyper = (ctypes.c_double * 9)(0.9, 0.999, 1e-8, sched.factor, sched.patience, sched.threshold, sched.min_lr,
                                          sched.eps, sched.max_n_plateaus)
            call("xvr_reg_update", ptr(reg.rotation), ptr(reg.translation), ptr(reg.rotation.grad.contiguous()),
                 ptr(reg.translation.grad.contiguous()), reg.rotation.numel(), ptr(state["exp_avg"][0]),
                 ptr(state["exp_avg_sq"][0]), ptr(state["exp_avg"][1]), ptr(state["exp_avg_sq"][1]), ptr(state["packed"]),
                 ptr(loss.detach()), ptr(log["rows"]), ptr(log["count"]), log["rows"].shape[0], hyper, stream())
            return
        with torch.no_grad():
            adam_maximize_([reg.rotation, reg.translation], [reg.rotation.grad, reg.translation.grad], state,
                           sched.lrs, active=sched.active)
            was_active = sched.active.clone()
            sched.step(loss.detach())
            # row `count` of the log <- (similarity before the update, parameters after it, next learning rates)
            idx = log["count"].long().reshape(1)
            row = torch.cat([loss.detach().reshape(1), reg.rotation.reshape(-1), reg.translation.reshape(-1),
                             torch.stack(sched.lrs).float()])
            keep = log["rows"].index_select(0, idx)[0]
            log["rows"].index_copy_(0, idx, torch.where(was_active > 0, row, keep)[None])
            log["count"].add_(was_active.float())

    # ------------------------------------------------------------------ public API
    def run(self, gt, init_pose, intrinsics=None, verbose=False):
        if len(init_pose) != 1 or gt.shape[0] != 1:
            # the fused update kernel (csrc/regstep.cu) and the trajectory rows are laid out for one pose, as the
            # reference's loop is (one X-ray per `xvr register` call, base.py:198-292)
            raise ValueError(f"Registrar.run registers ONE X-ray to one pose; got {gt.shape[0]} image(s) and "
                             f"{len(init_pose)} initial pose(s)")
        device = self.drr.device
        self.sim2.to(device)
        if intrinsics is not None:
            self.drr.set_intrinsics_(**intrinsics)
        height = gt.shape[-2]
        scales = parse_scales(self.scales, self.crop, height)
        rot, xyz = init_pose.convert(self.parameterization, self.convention)
        reg = Registration(self.drr, rot.to(device), xyz.to(device), self.parameterization, self.convention)
        n_rot = reg.rotation.numel()

        stage_rows, stage_times, stage_counts = [], [], []
        step_size_scalar = 1.0
        for stage, (scale, n_itr) in enumerate(zip(scales, self.n_itrs), start=1):
            reg.drr.rescale_detector_(scale)
            transform = XrayTransforms(reg.drr.detector.height, reg.drr.detector.width, equalize=self.equalize)
            img = transform(gt.to(device))

            step_size_scalar *= 2 ** (stage - 1)
            # {Adam step, best, num_bad, smallest lr seen, n_plateaus, active, lr_rot, lr_xyz}: one buffer, viewed by
            # the tensor-op implementations below and owned by xvr_reg_update when fused_update is on
            packed = torch.zeros(8, device=device, dtype=torch.float64)
            packed[6], packed[7] = self.lr_rot / step_size_scalar, self.lr_xyz / step_size_scalar
            lrs = [packed[6], packed[7]]
            sched = PlateauScheduler(lrs, factor=0.1, patience=self.patience, threshold=self.threshold,
                                     max_n_plateaus=self.max_n_plateaus, storage=packed[1:6])
            state = {"step": packed[0], "packed": packed,
                     "exp_avg": [torch.zeros_like(reg.rotation), torch.zeros_like(reg.translation)],
                     "exp_avg_sq": [torch.zeros_like(reg.rotation), torch.zeros_like(reg.translation)]}
            log = {"rows": torch.zeros(max(n_itr, 1), 1 + n_rot + 3 + 2, device=device),
                   "count": torch.zeros((), device=device)}

            graph = None
            if self.use_cuda_graph and n_itr > 0:
                graph = self._capture(reg, transform, img, state, sched, log)

            done, chunk_times = 0, []
            torch.cuda.synchronize()
            while done < n_itr:
                n = min(self.poll_every, n_itr - done)
                t0 = time.time()
                for _ in range(n):
                    if graph is not None:
                        graph.replay()
                    else:
                        self._iteration(reg, transform, img, state, sched, log)
                still_active = bool(sched.active.item() > 0)  # the only host round trip, once per chunk
                chunk_times.append((time.time() - t0, n))
                done += n
                if not still_active:
                    break
            count = int(log["count"].item())
            per_iter = []
            for dt, n in chunk_times:
                per_iter += [dt / n] * n
            stage_rows.append(log["rows"][:count].cpu())
            stage_times.append(per_iter[:count])
            stage_counts.append(count)
            if verbose:
                last = stage_rows[-1][-1, 0].item() if count else float("nan")
                print(f"stage {stage}: {count} iterations, similarity {last:.4f}")

        with torch.no_grad():
            final = float(self._score(transform, img, reg()))
            pose = reg.pose
        rows = torch.cat(stage_rows) if stage_rows else torch.zeros(0, 1 + n_rot + 5)
        traj = convert(rows[:, 1:1 + n_rot], rows[:, 1 + n_rot:4 + n_rot], parameterization=self.parameterization,
                       convention=self.convention).convert("euler_angles", "ZXY")
        init_row = torch.cat(init_pose.convert("euler_angles", "ZXY"), dim=-1).reshape(-1).tolist()
        params = [init_row] + torch.cat(traj, dim=-1).tolist()
        nccs = rows[:, 0].tolist() + [final]
        alphas = [[self.lr_rot, self.lr_xyz]] + rows[:, -2:].tolist()
        times = [0.0] + [t for st in stage_times for t in st]
        return pose, dict(params=params, nccs=nccs, times=times, alphas=alphas, runtime=sum(times),
                          n_itrs=stage_counts)

    # ------------------------------------------------------------------ results on disk
    COLUMNS = ("r1", "r2", "r3", "tx", "ty", "tz", "ncc", "times", "lr_rot", "lr_xyz")

    def register(self, gt, init_pose, intrinsics, outpath, name="xray", verbose=False):
        """``_RegistrarBase.__call__`` (base.py:294-339) for an X-ray that is already a tensor: run, render the
        initial / final DRRs if ``saveimg``, write ``<outpath>/<name>/parameters.pt``.  ``intrinsics`` are the
        X-ray's own (sdd, height, width, delx, dely, x0, y0 as read from its DICOM); like the reference's ``run``
        (base.py:141-149) the principal-point x offset is negated before it reaches the detector."""
        from pathlib import Path  # noqa: PLC0415

        applied = dict(intrinsics)
        if "x0" in applied:
            applied["x0"] = -applied["x0"]
        final_pose, info = self.run(gt, init_pose, applied, verbose=verbose)
        savepath = Path(outpath) / name
        savepath.mkdir(parents=True, exist_ok=True)
        init_img = final_img = None
        if self.saveimg:
            with torch.no_grad():
                init_img = self.drr(init_pose.to(self.drr.device)).cpu()
                final_img = self.drr(final_pose).cpu()
        trajectory = self.trajectory(info)
        self.save(savepath, gt, init_img, final_img, name, applied, init_pose.matrix.detach().cpu(),
                  final_pose.matrix.detach().cpu(), dict(pf_to_af=None, runtime=info["runtime"], trajectory=trajectory))
        return final_pose, info

    def trajectory(self, info):
        """The reference's ``_make_csv`` (base.py:410-423): one row per iteration with COLUMNS; a pandas DataFrame
        when pandas is importable (as in the reference), else a dict of column -> list."""
        import numpy as np  # noqa: PLC0415

        table = np.concatenate([np.asarray(info["params"], dtype=np.float64),
                                np.asarray(info["nccs"], dtype=np.float64)[:, None],
                                np.asarray(info["times"], dtype=np.float64)[:, None],
                                np.asarray(info["alphas"], dtype=np.float64)], axis=1)
        try:
            import pandas as pd  # noqa: PLC0415

            return pd.DataFrame(table, columns=list(self.COLUMNS))
        except ImportError:
            return {c: table[:, i].tolist() for i, c in enumerate(self.COLUMNS)}

    def save(self, savepath, gt, init_img, final_img, i2d, intrinsics, init_pose, final_pose, kwargs):
        """Write ``parameters.pt`` with the keys of the reference's ``_RegistrarBase.save`` (base.py:341-399), so that
        downstream scripts that read registration results (``torch.load(".../parameters.pt")["final_pose"]`` ...) work
        unchanged; PNGs of the target / initial / final images with ``saveimg``."""
        from pathlib import Path  # noqa: PLC0415

        pv = self.provenance
        as_path = lambda p: Path(p).resolve() if p is not None else None  # noqa: E731
        renderer = pv.get("renderer", type(self.drr.renderer).__name__.lower())
        parameters = {
            "drr": {
                "volume": as_path(pv.get("volume")),
                "mask": as_path(pv.get("mask")),
                "labels": pv.get("labels"),
                "orientation": pv.get("orientation", "AP"),
                **intrinsics,
                "reverse_x_axis": pv.get("reverse_x_axis", self.drr.detector.reverse_x_axis),
                "renderer": renderer,
                "read_kwargs": pv.get("read_kwargs", {}),
                "drr_kwargs": pv.get("drr_kwargs", {}),
            },
            "xray": {
                "filename": as_path(i2d) if pv.get("xray_is_file", False) else i2d,
                "crop": self.crop,
                "subtract_background": pv.get("subtract_background", False),
                "linearize": pv.get("linearize", True),
                "reducefn": pv.get("reducefn", "max"),
            },
            "optimization": {
                "equalize": self.equalize,
                "init_only": False,
                "scales": ",".join(self.scales),
                "n_itrs": ",".join(str(n) for n in self.n_itrs),
                "parameterization": self.parameterization,
                "convention": self.convention,
                "lr_rot": self.lr_rot,
                "lr_xyz": self.lr_xyz,
                "patience": self.patience,
                "max_n_plateaus": self.max_n_plateaus,
            },
            "init_pose": init_pose,
            "final_pose": final_pose,
            **pv.get("save_kwargs", {}),
            **kwargs,
        }
        torch.save(parameters, f"{savepath}/parameters.pt")
        if self.saveimg:
            from torchvision.utils import save_image  # noqa: PLC0415

            save_image(gt, f"{savepath}/gt.png", normalize=True)
            save_image(init_img, f"{savepath}/init_img.png", normalize=True)
            if final_img is not None:
                save_image(final_img, f"{savepath}/final_img.png", normalize=True)
        return parameters

    # ------------------------------------------------------------------ CUDA graph
    def _capture(self, reg, transform, img, state, sched, log):
        """Capture one iteration.  Everything it touches is updated in place, so replaying the graph advances the
        optimisation; the state is snapshotted and restored around the warm-up and capture runs."""
        tensors = [reg.rotation, reg.translation, state["step"], *state["exp_avg"], *state["exp_avg_sq"],
                   *sched.state(), log["rows"], log["count"]]
        snapshot = [t.detach().clone() for t in tensors]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: texture upload, allocator pools, lazy initialisation
                self._iteration(reg, transform, img, state, sched, log)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._iteration(reg, transform, img, state, sched, log)
        with torch.no_grad():
            for dst, src in zip(tensors, snapshot):
                dst.copy_(src)
        reg.rotation.grad = None
        reg.translation.grad = None
        return graph


def make_target(drr, rot, xyz, parameterization="euler_angles", convention="ZXY"):
    """Synthetic ground-truth X-ray: the DRR at a known pose (BASELINE config 3)."""
    with torch.no_grad():
        return drr(convert(rot, xyz, parameterization=parameterization, convention=convention))
