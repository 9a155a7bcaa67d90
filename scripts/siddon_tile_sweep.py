"""Siddon forward at config-5 geometry (768^3, 512^2) for a sweep of warp/CTA detector tile shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, xvr_b200
from xvr_b200.data import read, synthetic_ct

dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hu, _, aff = synthetic_ct(768, device=dev)
drr = xvr_b200.DRR(read(hu, affine=aff), bench.SDD, 512, bench.DELX / 2, renderer="siddon", reverse_x_axis=False).to(dev)
del hu
rot, xyz = (t.to(dev) for t in bench.pose_batch(B, 1))
pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
src, tgt = drr.detector(pose, None)
raylen = (tgt - src).norm(dim=-1).unsqueeze(1).contiguous()
src, tgt = drr.affine_inverse(src).contiguous(), drr.affine_inverse(tgt).contiguous()
res = {}
for lw, cw in [(0, 3), (0, 0), (1, 3), (2, 3), (2, 4), (3, 3), (3, 5), (1, 4), (5, 8), (4, 4)]:
    if cw < lw or (8 - cw) < (5 - lw):
        continue
    os.environ["XVR_B200_TILE"] = f"{lw},{cw}"
    with torch.no_grad():
        for _ in range(2):
            drr.renderer(drr.density, src, tgt, raylen)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            drr.renderer(drr.density, src, tgt, raylen)
        e1.record()
        torch.cuda.synchronize()
    res[f"{lw},{cw}"] = round(e0.elapsed_time(e1) / 3, 3)
print(json.dumps({"B": B, "ms": res}))
