"""How much margin does the production tolerance of the Siddon fast-index certificate have?  Traces config-5 rays
(768^3, 512^2) with the tolerance scaled down and counts voxel indices that differ from the always-exact path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, xvr_b200
from xvr_b200._lib import call, options, opts_word, ptr, stream
from xvr_b200.data import read, synthetic_ct

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
hu, _, aff = synthetic_ct(n, device=dev)
drr = xvr_b200.DRR(read(hu, affine=aff), bench.SDD, 512, bench.DELX / 2, renderer="siddon", reverse_x_axis=False).to(dev)
del hu
out = {}
for seed in (0, 1):
    rot, xyz = (t.to(dev) for t in bench.pose_batch(1, seed))
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    src, tgt = drr.detector(pose, None)
    src, tgt = drr.affine_inverse(src).contiguous(), drr.affine_inverse(tgt).contiguous()
    B, N, M = 1, tgt.shape[1], 3 * n + 8

    def trace(scale):
        idx = torch.full((B, N, M), -2, dtype=torch.int32, device=dev)
        seg = torch.zeros(B, N, M, device=dev)
        cnt = torch.zeros(B, N, dtype=torch.int32, device=dev)
        with options(siddon_tol=scale):  # per-call option word (include/xvr_b200.h XVR_OPT_SIDDON_TOL)
            call("xvr_siddon_trace", ptr(drr.density), None, *drr.density.shape, ptr(src), ptr(tgt), B, N, 0.5, 1e-8, M, ptr(idx),
                 ptr(seg), ptr(cnt), opts_word(), stream())
        return idx, cnt

    exact, cnt = trace("exact")
    for scale in ("production", 0.5, 0.25, 0.125):
        idx, _ = trace(scale)
        out[f"seed{seed}/scale{scale}"] = int((idx != exact).sum().item())
    out[f"seed{seed}/segments"] = int(cnt.sum().item())
print(json.dumps(out))
