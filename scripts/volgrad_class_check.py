"""CPU check of the race-freedom argument of volume_grad_brick_kernel (csrc/volgrad.cu): for every
brick and pose, rays that the kernel would put in the same colour class (S pixels apart, S from the kernel's own
formula) must never touch a common voxel of the brick.  Prints the S histogram, the number of same-class ray pairs
checked, the number of clashes (must be 0) and the smallest L-infinity distance between their samples (must be > 2).
Geometry: 64^3 volume of 4 mm voxels, 32^2 detector of 8.7 mm pixels = the pixel/voxel ratio of BASELINE config 2."""
import sys, math, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, xvr_b200
from tests.golden.make_golden import scene
def run(N, H, W, NP, VGB, rot, xyz):
    hu, _, affine = scene(N)
    affinv = torch.as_tensor(np.linalg.inv(affine)).double()
    sdd = 1020.0; delx = 1.08821875 * 256 / H; dely = delx
    det = xvr_b200.Detector(sdd, H, W, delx, dely, 0.0, 0.0, oracle.REORIENT["AP"], reverse_x_axis=False)
    o, u, v = (np.array(t) for t in det.pixel_basis())
    pose = oracle.pose_from_params(rot.double(), xyz.double(), "euler_angles", "ZXY")
    cam2vox = (affinv @ (pose @ oracle.REORIENT["AP"].double())).numpy(); vox2cam = np.linalg.inv(cam2vox)
    D = np.array([N, N, N]); eps = 1e-8
    k = np.arange(NP); st = 1.0 / (NP - 1)
    lin = np.where(k < NP // 2, st * k, 1.0 - st * (NP - 1 - k))
    offs = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    stats = dict(S={}, checked=0, clashes=0, minsep=9e9)
    nb = (N + VGB - 1) // VGB
    for b in range(len(rot)):
      G = cam2vox[b]; Gi = vox2cam[b]; s = G[:3, 3]
      # all rays once
      ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
      c = o[None, None] + ii[..., None] * u + jj[..., None] * v
      d = (c @ G[:3, :3].T) + eps
      a0 = (0 - s) / d; a1 = (D - 1 - s) / d
      amin = np.maximum(np.minimum(a0, a1).max(-1), 0); amax = np.minimum(np.maximum(a0, a1).min(-1), 1)
      for bx in range(nb):
        for by in range(nb):
          for bz in range(nb):
            lo = np.array([bx, by, bz]) * VGB; elo, ehi = lo - 1.0, lo + float(VGB)
            P = np.array([[ehi[0] if cc & 1 else elo[0], ehi[1] if cc & 2 else elo[1], ehi[2] if cc & 4 else elo[2]] for cc in range(8)])
            Q = P @ Gi[:3, :3].T + Gi[:3, 3]; m = sdd / Q[:, 2]
            cj = (Q[:, 0] * m - o[0]) / v[0]; ci = (Q[:, 1] * m - o[1]) / u[1]; dmin = Q[:, 2].min()
            j0, j1 = max(0, math.floor(cj.min())), min(W - 1, math.ceil(cj.max()))
            i0, i1 = max(0, math.floor(ci.min())), min(H - 1, math.ceil(ci.max()))
            if j0 > j1 or i0 > i1: continue
            sp2 = max((Gi[:3, a] ** 2).sum() for a in range(3))
            xmax = max(abs(o[0]), abs(o[0] + v[0] * (W - 1))); ymax = max(abs(o[1]), abs(o[1] + u[1] * (H - 1)))
            cos2 = sdd * sdd / (sdd * sdd + xmax * xmax + ymax * ymax)
            sep = cos2 * (dmin / sdd) * min(abs(u[1]), abs(v[0])) / math.sqrt(sp2) * 0.57735027
            S = max(1, math.ceil(2.05 / max(sep, 1e-6))); stats["S"][S] = stats["S"].get(S, 0) + 1
            foot = {}
            for i in range(i0, i1 + 1):
              for j in range(j0, j1 + 1):
                if not amax[i, j] > amin[i, j]: continue
                x = (lin * (amax[i, j] - amin[i, j]) + amin[i, j])[:, None] * d[i, j] + s
                inside = ((x > elo) & (x < ehi)).all(1)
                if not inside.any(): continue
                f0 = np.floor(x[inside]).astype(int) - lo
                cells = (f0[:, None, :] + offs[None]).reshape(-1, 3)
                cells = cells[((cells >= 0) & (cells < VGB)).all(1)]
                if len(cells): foot[(i, j)] = (set(map(tuple, cells)), x[inside])
            keys = list(foot)
            for a in range(len(keys)):
              for bb in range(a + 1, len(keys)):
                (ia, ja), (ib, jb) = keys[a], keys[bb]
                if (ia - ib) % S == 0 and (ja - jb) % S == 0:
                    stats["checked"] += 1
                    if foot[keys[a]][0] & foot[keys[bb]][0]: stats["clashes"] += 1
                    xa, xb = foot[keys[a]][1], foot[keys[bb]][1]
                    dd = np.abs(xa[:, None, :] - xb[None, :, :]).max(-1).min()
                    stats["minsep"] = min(stats["minsep"], dd)
    return stats
rot = torch.tensor([[0.2, -0.3, 0.1], [-0.7, 0.6, -0.2]]); xyz = torch.tensor([[10., 780., -15.], [-45., 700., 40.]])
print(run(64, 32, 32, 125, 16, rot, xyz))
