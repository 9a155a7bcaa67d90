"""Experimental brick-local dL/dvolume kernel (xvr_set_volgrad_version(2)) vs the default gather kernel:
agreement (with the worst voxels listed) and time.  usage: volgrad_v2_probe.py [n_vol det batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, xvr_b200
from xvr_b200._lib import call

n, det, B = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 256, 8)
drr = bench.build_scene(torch.device("cuda"), n, det)
rot, xyz = (t.cuda() for t in bench.pose_batch(B, 0))
vol = drr.density.detach().clone().requires_grad_()
drr.density = vol
img = drr(xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY"))
g = torch.rand_like(img)
out = {"n": n, "det": det, "B": B}
grads = {}
for version in (1, 2):
    call("xvr_set_volgrad_version", version)
    vol.grad = None
    img.backward(g, retain_graph=True)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vol.grad = None
    e0.record()
    img.backward(g, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    grads[version] = vol.grad.clone()
    out[f"v{version}_ms"] = e0.elapsed_time(e1)
call("xvr_set_volgrad_version", 1)
diff = grads[2] - grads[1]
out["rel_l2"] = (diff.norm() / grads[1].norm()).item()
out["max_abs"] = grads[1].abs().max().item()
out["n_bad"] = int((diff.abs() > 1e-4 * out["max_abs"]).sum())
top = diff.abs().flatten().topk(12).indices
worst = []
for i in top.tolist():
    x, y, z = i // (n * n), (i // n) % n, i % n
    worst.append([x, y, z, x % 16, y % 16, z % 16, round(grads[1][x, y, z].item(), 5), round(grads[2][x, y, z].item(), 5)])
out["worst_xyz_local_v1_v2"] = worst
print(json.dumps(out))
