"""Offline model of a TMA-staged, slab-marching trilinear forward kernel at BASELINE config 2 (512^3, 256^2, 500
samples per ray, bench poses) -- the design north_star names and DESIGN.md 5.1 has so far only argued about.

A CTA owns a TH x TW detector tile of one pose and marches its rays slab by slab along the tile's dominant volume
axis.  For a slab of K voxel layers it stages the box (K + 1 layers) x (extent of the tile's samples on the other two
axes, + 1) into shared memory (one cp.async.bulk.tensor.3d; out-of-range elements arrive as zeros, which IS
grid_sample's zero padding), then every thread interpolates its samples of that slab from 8 shared-memory loads.

The model reports, per tile shape and slab thickness:
  fill GB/launch    bytes staged per launch of 116 poses (to compare with the texture path's measured L2->L1 fill:
                    2.47 G sectors = 79 GB, profiles/r1_trilinear_fwd_ncu_full.md)
  box KB            largest staged box (x2 for double buffering -> CTAs per SM)
  reuse             corner reads served per staged voxel
  conflict          mean shared-memory wavefronts per warp-wide corner load (distinct words per bank, max over
                    banks), for the 32 lanes of a warp taking the same sample index k on 32 adjacent rays
No GPU needed; run:  python scripts/sim_brick_staging.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from bench import DELX, SDD, pose_batch

N, DET, NP, BATCH = 512, 256, 500, 116
sp = 256.0 / N
aff = torch.diag(torch.tensor([sp, sp, sp, 1.0]))
aff[:3, 3] = -sp * (N - 1) / 2
affinv = torch.linalg.inv(aff)[None]
P = 6  # poses modelled
rot, xyz = pose_batch(BATCH, 0)
pose = oracle.pose_from_params(rot[:P], xyz[:P], "euler_angles", "ZXY")
src, tgt = oracle.detector_rays(pose, oracle.REORIENT["AP"], DET, DET, DELX, DELX, 0.0, 0.0, SDD, False)
src, tgt = oracle.apply(affinv, src), oracle.apply(affinv, tgt)
amin, amax = oracle.alpha_minmax(src, tgt, torch.tensor([N - 1.0] * 3), 1e-8)
src = src.numpy().astype(np.float64)[:, 0]                       # (P,3)
tgt = tgt.view(P, DET, DET, 3).numpy().astype(np.float64)        # (P,H,W,3)
amin = amin.view(P, DET, DET).numpy().astype(np.float64)
amax = amax.view(P, DET, DET).numpy().astype(np.float64)
u = np.linspace(0.0, 1.0, NP)


def tile_samples(b, i0, j0, th, tw):
    """Sample positions (th, tw, NP, 3) of a detector tile, NaN for rays that miss the volume."""
    d = tgt[b, i0:i0 + th, j0:j0 + tw] - src[b]
    a0, a1 = amin[b, i0:i0 + th, j0:j0 + tw], amax[b, i0:i0 + th, j0:j0 + tw]
    alpha = a0[..., None] + (a1 - a0)[..., None] * u
    x = src[b] + alpha[..., None] * d[:, :, None, :]
    x[~(a1 > a0)] = np.nan
    return x


def conflicts(words):
    """words (G,32) int64 shared-memory word addresses (-1 = inactive lane) -> mean wavefronts per request."""
    out = []
    for w in words:
        w = np.unique(w[w >= 0])
        if len(w) == 0:
            continue
        out.append(np.bincount(w % 32, minlength=32).max())
    return float(np.mean(out)) if out else float("nan")


def model(th, tw, K, warp_h, warp_w, n_tiles=24, seed=0):
    g = np.random.default_rng(seed)
    fill = corner_reads = staged = 0.0
    max_box = 0
    conf = []
    n = 0
    for b in range(P):
        for _ in range(n_tiles // P):
            i0 = int(g.integers(0, DET // th)) * th
            j0 = int(g.integers(0, DET // tw)) * tw
            x = tile_samples(b, i0, j0, th, tw)
            ok = ~np.isnan(x[..., 0])
            if ok.sum() == 0:
                continue
            n += 1
            centre = tgt[b, i0 + th // 2, j0 + tw // 2] - src[b]
            A = int(np.abs(centre).argmax())
            o1, o2 = [a for a in range(3) if a != A]
            ix = np.floor(x)
            slab = np.where(ok, ix[..., A] // K, -1).astype(np.int64)
            for s in np.unique(slab[slab >= 0]):
                m = slab == s
                lo1, hi1 = ix[..., o1][m].min(), ix[..., o1][m].max() + 1
                lo2, hi2 = ix[..., o2][m].min(), ix[..., o2][m].max() + 1
                e1, e2 = int(hi1 - lo1 + 1), int(hi2 - lo2 + 1)
                box = (K + 1) * e1 * e2
                max_box = max(max_box, box * 4)
                fill += box * 4
                staged += box
                corner_reads += 8 * m.sum()
                # bank conflicts of one warp-wide corner load: lanes = warp_h x warp_w adjacent rays, same k
                # shared layout: [layer along A][o1][o2 padded to odd]
                p2 = e2 | 1
                for _ in range(3):
                    wi = int(g.integers(0, th // warp_h)) * warp_h
                    wj = int(g.integers(0, tw // warp_w)) * warp_w
                    ks = np.where(m[wi:wi + warp_h, wj:wj + warp_w].any(axis=(0, 1)))[0]
                    if len(ks) == 0:
                        continue
                    k = int(g.choice(ks))
                    lane_ok = m[wi:wi + warp_h, wj:wj + warp_w, k].reshape(-1)
                    q = ix[wi:wi + warp_h, wj:wj + warp_w, k].reshape(-1, 3)
                    word = ((q[:, A] - s * K) * e1 + (q[:, o1] - lo1)) * p2 + (q[:, o2] - lo2)
                    word = np.where(lane_ok, word, -1).astype(np.int64)
                    conf.append(word)
    per_tile = fill / n
    tiles_per_drr = (DET // th) * (DET // tw)
    return dict(fill_gb=per_tile * tiles_per_drr * BATCH / 1e9, box_kb=max_box / 1024.0,
                reuse=corner_reads / staged, conflict=conflicts(np.array(conf)))


if __name__ == "__main__":
    print("tile  warp  K   fill GB/launch  largest box KB  reuse  wavefronts/LDS")
    for (th, tw), (wh, ww) in (((16, 16), (2, 16)), ((16, 16), (16, 2)), ((32, 8), (32, 1)), ((8, 32), (1, 32)),
                               ((32, 32), (2, 16))):
        for K in (4, 8, 16):
            r = model(th, tw, K, wh, ww)
            print(f"{th}x{tw:<3} {wh}x{ww:<3} {K:<3} {r['fill_gb']:>12.1f} {r['box_kb']:>15.1f} {r['reuse']:>7.2f} "
                  f"{r['conflict']:>10.2f}")
