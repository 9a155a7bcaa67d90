"""Offline model of L1 behaviour: distinct 128-byte lines touched per warp-wide corner load for candidate
thread->(ray, sample) mappings of the trilinear march (512^3 volume, 256^2 detector, bench poses)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import oracle
from bench import pose_batch, SDD, DELX

N, DET, NP = 512, 256, 500
sp = 256.0 / N
aff = torch.diag(torch.tensor([sp, sp, sp, 1.0])); aff[:3, 3] = -sp * (N - 1) / 2
affinv = torch.linalg.inv(aff)[None]
rot, xyz = pose_batch(116, 0)
P = 12
pose = oracle.pose_from_params(rot[:P], xyz[:P], "euler_angles", "ZXY")
src, tgt = oracle.detector_rays(pose, oracle.REORIENT["AP"], DET, DET, DELX, DELX, 0., 0., SDD, False)
src, tgt = oracle.apply(affinv, src), oracle.apply(affinv, tgt)
dims = torch.tensor([N - 1.0] * 3)
amin, amax = oracle.alpha_minmax(src, tgt, dims, 1e-8)
tgt = tgt.view(P, DET, DET, 3); amin = amin.view(P, DET, DET); amax = amax.view(P, DET, DET)
d = tgt - src.view(P, 1, 1, 3)

def lines_for(b, ii, jj, kk):
    """ii,jj,kk: (G,32) integer tensors -> mean distinct lines / sectors per request over the 8 corners."""
    a = amin[b, ii, jj] + (kk.float() / (NP - 1)) * (amax[b, ii, jj] - amin[b, ii, jj])
    p = src[b, 0] + a[..., None] * d[b, ii, jj]
    valid = (amax[b, ii, jj] > amin[b, ii, jj])
    ip = p.floor().long().clamp(0, N - 2)
    tot_l = tot_s = 0.0
    for c in range(8):
        q = ip + torch.tensor([(c >> 2) & 1, (c >> 1) & 1, c & 1])
        addr = ((q[..., 0] * N + q[..., 1]) * N + q[..., 2]) * 4
        addr = torch.where(valid, addr, torch.full_like(addr, -1))
        for shift, acc in ((7, "l"), (5, "s")):
            u = addr >> shift
            u = torch.where(valid, u, torch.full_like(u, -1))
            su = u.sort(dim=1).values
            cnt = 1 + (su[:, 1:] != su[:, :-1]).sum(1) - (~valid).any(1).long()
            if acc == "l": tot_l += cnt.float().mean().item()
            else: tot_s += cnt.float().mean().item()
    return tot_l / 8, tot_s / 8

g = torch.Generator().manual_seed(0)
G = 400
def sample_base(h, w):
    i0 = (torch.randint(0, DET // h, (G,), generator=g) * h)
    j0 = (torch.randint(0, DET // w, (G,), generator=g) * w)
    k0 = torch.randint(60, NP - 60, (G,), generator=g)
    return i0, j0, k0

res = {}
for b in range(P):
    row = {}
    for name, (h, w) in {"8x4": (4, 8), "32x1": (32, 1), "1x32": (1, 32), "16x2": (16, 2), "2x16": (2, 16)}.items():
        i0, j0, k0 = sample_base(h, w)
        l = torch.arange(32)
        ii = i0[:, None] + (l // w)[None]; jj = j0[:, None] + (l % w)[None]; kk = k0[:, None].expand(-1, 32)
        row[name] = lines_for(b, ii, jj, kk)
    # skewed: lanes along i (or j), per-lane column shear and k-skew chosen so that the lanes line up along z
    for name, along_i in (("skew_i", True), ("skew_j", False)):
        i0, j0, k0 = sample_base(32, 32)
        # linearise p(i,j,k) at the tile centre
        ic, jc = i0 + 16, j0 + 16
        def pos(i, j, k):
            a = amin[b, i, j] + (k.float() / (NP - 1)) * (amax[b, i, j] - amin[b, i, j])
            return src[b, 0] + a[..., None] * d[b, i, j]
        pi = pos((ic + 1).clamp(max=DET-1), jc, k0) - pos(ic, jc, k0)
        pj = pos(ic, (jc + 1).clamp(max=DET-1), k0) - pos(ic, jc, k0)
        pk = pos(ic, jc, k0 + 1) - pos(ic, jc, k0)
        M = torch.stack([pi, pj, pk], -1)  # columns
        ez = torch.tensor([0., 0., 1.]).expand(G, 3)
        sol = torch.linalg.solve(M, ez[..., None])[..., 0]  # (di,dj,dk) per unit z
        lead = sol[:, 0] if along_i else sol[:, 1]
        sol = sol / lead[:, None]
        l = torch.arange(32).float()[None]
        if along_i:
            ii = i0[:, None] + l.long(); jj = (j0[:, None] + 16 + (l * sol[:, 1:2]).round().long()) % DET
        else:
            jj = j0[:, None] + l.long(); ii = (i0[:, None] + 16 + (l * sol[:, 0:1]).round().long()) % DET
        kk = (k0[:, None] + (l * sol[:, 2:3]).round().long()).clamp(0, NP - 1)
        row[name] = lines_for(b, ii, jj, kk)
        row[name + "_zstep"] = (1.0 / lead.abs()).median().item()
    res[b] = row
    print(b, " ".join(f"{k}={v[0]:.1f}/{v[1]:.1f}" if isinstance(v, tuple) else f"{k}={v:.2f}" for k, v in row.items()))
