"""BASELINE config 3: xvr register on a 512^3 CT, 256x256 target, mNCC+GradNCC + Adam, 200 iterations, 1xB200."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, xvr_b200
from xvr_b200.registrar import Registrar

# usage: bench_register.py [volume size] [--unfused-similarity]   (default: xvr_regsim)
fused = "--unfused-similarity" not in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
vol = int(args[0]) if args else 512
drr = bench.build_scene(torch.device("cuda"), vol, 256)
rot0 = torch.tensor([[0.20, -0.10, 0.05]], device="cuda"); xyz0 = torch.tensor([[5.0, 800.0, -10.0]], device="cuda")
with torch.no_grad():
    gt = drr(xvr_b200.convert(rot0, xyz0, parameterization="euler_angles", convention="ZXY"))
d = torch.deg2rad(torch.tensor([[5.0, 5.0, 5.0]], device="cuda")); t = torch.tensor([[10.0, 10.0, 10.0]], device="cuda")
out = {}
for graph in (True, False):
    init = xvr_b200.convert(rot0 + d, xyz0 + t, parameterization="euler_angles", convention="ZXY")
    reg = Registrar(drr, scales="1", n_itrs="200", max_n_plateaus=10**6, use_cuda_graph=graph, poll_every=50,
                    fused_similarity=fused)
    reg.run(gt, init)  # warm-up (graph capture, allocator)
    torch.cuda.synchronize(); t0 = time.time()
    pose, info = reg.run(gt, init)
    torch.cuda.synchronize(); dt = time.time() - t0
    rot, xyz = pose.convert("euler_angles", "ZXY")
    out["graph" if graph else "eager"] = {"iters": info["n_itrs"][0], "wall_s": dt, "loop_ms_per_iter": 1e3 * info["runtime"] / max(1, info["n_itrs"][0]),
        "final_ncc": info["nccs"][-1], "rot_err_deg": torch.rad2deg((rot - rot0).abs().max()).item(), "xyz_err_mm": (xyz - xyz0).abs().max().item()}
print(json.dumps({"workload": f"register {vol}^3 CT, 256x256, B=1, 200 iters, scales=1", "fused_similarity": fused, **out}))
