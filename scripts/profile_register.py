import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, xvr_b200
from xvr_b200.registrar import Registrar
from torch.profiler import profile, ProfilerActivity
drr = bench.build_scene(torch.device("cuda"), 512, 256)
rot0 = torch.tensor([[0.20, -0.10, 0.05]], device="cuda"); xyz0 = torch.tensor([[5.0, 800.0, -10.0]], device="cuda")
with torch.no_grad():
    gt = drr(xvr_b200.convert(rot0, xyz0, parameterization="euler_angles", convention="ZXY"))
init = xvr_b200.convert(rot0 + 0.05, xyz0 + 8.0, parameterization="euler_angles", convention="ZXY")
graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
reg = Registrar(drr, scales="1", n_itrs="20", max_n_plateaus=10**6, use_cuda_graph=graph, poll_every=50)
reg.run(gt, init)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    reg.run(gt, init)
    torch.cuda.synchronize()
ka = prof.key_averages()
print(ka.table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
kern = [e for e in ka if e.device_type.name == "CUDA"]
print("GPU kernels per iteration:", sum(e.count for e in kern) / 20, " GPU time per iteration (us):", sum(e.device_time_total for e in kern) / 20)
