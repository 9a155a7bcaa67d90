set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
python - <<'PY'
import torch, time, xvr_b200, bench
drr = bench.build_scene(torch.device("cuda"))
rot, xyz = bench.pose_batch(116, 0)
rot, xyz = rot.cuda(), xyz.cuda()
vol = drr.density.detach().clone().requires_grad_()
drr.density = vol
gout = torch.rand(116,1,256,256, device="cuda")
for nb in (4, 16, 116):
    pose = xvr_b200.convert(rot[:nb], xyz[:nb], parameterization="euler_angles", convention="ZXY")
    img = drr(pose)
    torch.cuda.synchronize(); 
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); img.backward(gout[:nb]); e1.record(); torch.cuda.synchronize()
    print("poses", nb, "volume-grad backward ms", e0.elapsed_time(e1), "gvol norm", vol.grad.norm().item())
    vol.grad=None
PY
