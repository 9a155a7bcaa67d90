import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, xvr_b200
drr = bench.build_scene(torch.device("cuda"), 512, 256)
B = 8
rot, xyz = (t.cuda() for t in bench.pose_batch(B, 0))
vol = drr.density.detach().clone().requires_grad_()
drr.density = vol
img = drr(xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY"))
g = torch.rand_like(img)
for _ in range(2):
    vol.grad = None
    img.backward(g, retain_graph=True)
torch.cuda.synchronize()
