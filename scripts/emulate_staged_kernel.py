"""CPU transliteration of csrc/trilinear_staged.cu's control flow and index arithmetic (NOT of its fp32 rounding):
slab numbering, travel order, the per-slab box from the tile's corner rays (margins, clipping, 16-byte alignment,
capacity), the row-wise staging with zero fill, the local corner addressing and the in-box test.

For every tile it replays the kernel thread by thread in float64 and checks that
  * every sample of every ray is consumed exactly once and in ascending order,
  * a sample served from the staged buffer reads exactly the 8 voxels (or zero padding) a direct lookup reads,
  * the fraction served from the buffer is what the design expects (~all).
Written because the kernel could not be run when it was written (round 1's GPU budget was spent); it checks the
logic a GPU run would otherwise be the first to exercise.   python scripts/emulate_staged_kernel.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle

ST_T, ST_K, ST_CAP = 16, 4, 12288


def slab_of(cell):
    return (max(cell, -ST_K) + ST_K) // ST_K


def direct_corners(vol, ix, iy, iz):
    out = np.zeros(8)
    D = vol.shape
    for c in range(8):
        x, y, z = ix + (c >> 2 & 1), iy + (c >> 1 & 1), iz + (c & 1)
        if 0 <= x < D[0] and 0 <= y < D[1] and 0 <= z < D[2]:
            out[c] = vol[x, y, z]
    return out


def run_tile(vol, src, tgt, amin, amax, i0, j0, H, W, n_points, stats):
    D = vol.shape
    u = np.linspace(0.0, 1.0, n_points)
    ci, cj = min(i0 + ST_T // 2, H - 1), min(j0 + ST_T // 2, W - 1)
    cd = tgt[ci, cj] - src
    A = 1 if abs(cd[1]) > abs(cd[0]) else 0
    if abs(cd[2]) > abs(cd[A]):
        A = 2
    forward = cd[A] > 0
    O1, O2 = (1 if A == 0 else 0), (1 if A == 2 else 2)
    corners = [(min(i0 + ST_T - 1, H - 1) if l & 1 else i0, min(j0 + ST_T - 1, W - 1) if l & 2 else j0) for l in range(4)]

    threads = []
    t_first, t_last = None, None
    for tid in range(256):
        lane, warp = tid & 31, tid >> 5
        pi, pj = i0 + (lane & 15), j0 + warp * 2 + (lane >> 4)
        if not (pi < H and pj < W):
            continue
        d = tgt[pi, pj] - src
        span = amax[pi, pj] - amin[pi, pj]
        regular = span > 0 and d[A] != 0 and ((d[A] > 0) == forward)
        if not regular:
            stats["irregular_rays"] += 1
            continue
        pos = src + (amin[pi, pj] + u * span)[:, None] * d
        cells = np.floor(pos).astype(int)
        tk = np.array([slab_of(c) if forward else -slab_of(c) for c in cells[:, A]])
        assert (np.diff(tk) >= 0).all(), "slab index not monotone along the ray"
        threads.append(dict(pos=pos, cells=cells, tk=tk, k=0))
        t_first = tk[0] if t_first is None else min(t_first, tk[0])
        t_last = tk[-1] if t_last is None else max(t_last, tk[-1])
    if t_first is None:
        return

    for t in range(t_first, t_last + 1):
        slab = t if forward else -t
        loA = (slab - 1) * ST_K
        pts = []
        for (qi, qj) in corners:
            qd = tgt[qi, qj] - src
            for plane in (loA, loA + ST_K):
                al = (plane - src[A]) / qd[A]
                pts.append(src + al * qd)
        pts = np.array(pts)
        l1, h1 = int(np.floor(pts[:, O1].min())) - 1, int(np.floor(pts[:, O1].max())) + 2
        l2, h2 = int(np.floor(pts[:, O2].min())) - 1, int(np.floor(pts[:, O2].max())) + 2
        lo, hi = [0, 0, 0], [0, 0, 0]
        for a in range(3):
            lo[a] = max(loA if a == A else (l1 if a == O1 else l2), -1)
            hi[a] = min(loA + ST_K if a == A else (h1 if a == O1 else h2), D[a])
        lo[2] &= ~3
        hi[2] = ((hi[2] + 4) & ~3) - 1
        E = [hi[a] - lo[a] + 1 for a in range(3)]
        elems = max(E[0], 0) * max(E[1], 0) * max(E[2], 0)
        staged = 0 < elems <= ST_CAP and D[2] % 4 == 0
        stats["slabs"] += 1
        stats["staged_slabs"] += int(staged)
        box = None
        if staged:
            box = np.full(elems, np.nan)
            c_lo, c_hi = max(lo[2], 0), min(lo[2] + E[2], D[2])
            assert c_lo % 4 == 0 and c_hi % 4 == 0 and lo[2] % 4 == 0 and E[2] % 4 == 0
            for r in range(E[0] * E[1]):
                g0, g1 = lo[0] + r // E[1], lo[1] + r % E[1]
                inside = 0 <= g0 < D[0] and 0 <= g1 < D[1] and c_hi > c_lo
                z_lo, z_hi = (c_lo - lo[2], c_hi - lo[2]) if inside else (E[2], E[2])
                box[r * E[2]: r * E[2] + z_lo] = 0.0
                box[r * E[2] + z_hi: (r + 1) * E[2]] = 0.0
                if inside:
                    box[r * E[2] + z_lo: r * E[2] + z_hi] = vol[g0, g1, c_lo:c_hi]
            assert not np.isnan(box).any(), "staging left part of the buffer unwritten"
        for th in threads:
            while th["k"] < n_points:
                k = th["k"]
                if th["tk"][k] > t:
                    break
                ix, iy, iz = th["cells"][k]
                lx, ly, lz = ix - lo[0], iy - lo[1], iz - lo[2]
                ref = direct_corners(vol, ix, iy, iz)
                if staged and th["tk"][k] == t and 0 <= lx < E[0] - 1 and 0 <= ly < E[1] - 1 and 0 <= lz < E[2] - 1:
                    q = (lx * E[1] + ly) * E[2] + lz
                    sy, sx = E[2], E[1] * E[2]
                    got = np.array([box[q], box[q + 1], box[q + sy], box[q + sy + 1], box[q + sx], box[q + sx + 1],
                                    box[q + sx + sy], box[q + sx + sy + 1]])
                    assert np.array_equal(got, ref), (t, k, (ix, iy, iz), lo, E)
                    stats["shared"] += 1
                else:
                    stats["global"] += 1
                    if th["tk"][k] < t:
                        stats["late"] += 1
                th["k"] += 1
    for th in threads:
        assert th["k"] == n_points, "a ray was left with unconsumed samples"


def main():
    from tests._scene import pixel_size

    SDD = 1020.0
    cases = [
        dict(shape=(40, 48, 44), spacing=(4.0, 4.0, 4.0), H=32, W=32, rot=[[0.2, -0.1, 0.05], [-0.6, 0.5, 0.2]],
             xyz=[[5.0, 800.0, -10.0], [-20.0, 750.0, 15.0]]),
        dict(shape=(40, 64, 52), spacing=(2.0, 1.5, 2.5), H=24, W=40,
             rot=[[0.0, 0.0, 0.0], [1.2, 0.3, 0.0], [0.0, 1.5707964, 0.0], [0.0, 0.0, 0.0]],
             xyz=[[0.0, 800.0, 0.0], [0.0, 300.0, 0.0], [60.0, 500.0, 50.0], [10.0, 30.0, -5.0]]),
    ]
    for case in cases:
        shape, sp = case["shape"], case["spacing"]
        vol = np.random.default_rng(0).random(shape)
        aff = torch.diag(torch.tensor([*sp, 1.0]))
        aff[:3, 3] = -torch.tensor(sp) * (torch.tensor(shape, dtype=torch.float32) - 1) / 2
        affinv = torch.linalg.inv(aff)[None]
        H, W = case["H"], case["W"]
        pose = oracle.pose_from_params(torch.tensor(case["rot"]), torch.tensor(case["xyz"]), "euler_angles", "ZXY")
        s, t = oracle.detector_rays(pose, oracle.REORIENT["AP"], H, W, pixel_size(H), pixel_size(H), 0.0, 0.0, SDD, False)
        s, t = oracle.apply(affinv, s), oracle.apply(affinv, t)
        amin, amax = oracle.alpha_minmax(s, t, torch.tensor(shape, dtype=torch.float32) - 1, 1e-8)
        stats = dict(slabs=0, staged_slabs=0, shared=0, irregular_rays=0, late=0)
        stats["global"] = 0
        for b in range(len(case["rot"])):
            src = s[b, 0].double().numpy()
            tgt = t[b].view(H, W, 3).double().numpy()
            a0, a1 = amin[b].view(H, W).double().numpy(), amax[b].view(H, W).double().numpy()
            for i0 in range(0, H, ST_T):
                for j0 in range(0, W, ST_T):
                    run_tile(vol, src, tgt, a0, a1, i0, j0, H, W, 60, stats)
        print(case["shape"], stats, "served from the buffer: %.4f" % (stats["shared"] / max(1, stats["shared"] + stats["global"])))
        assert stats["late"] == 0


if __name__ == "__main__":
    main()
