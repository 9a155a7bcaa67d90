import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, xvr_b200
from xvr_b200.data import read
GOLD = torch.load("tests/golden/oracle_v1.pt", weights_only=False)
dev = "cuda"
sub = read(GOLD["hu"], GOLD["labels"], affine=GOLD["affine"].numpy(), center_volume=False)
d = GOLD["detector"]
drr = xvr_b200.DRR(sub, d["sdd"], d["height"], d["delx"], d["width"], d["dely"], d["x0"], d["y0"], reverse_x_axis=d["reverse_x_axis"], renderer="siddon").to(dev)
pose = xvr_b200.convert(GOLD["rot"].to(dev), GOLD["xyz"].to(dev), parameterization="euler_angles", convention="ZXY")
src, tgt = drr.detector(pose, None)
raylen = (tgt - src).norm(dim=-1).unsqueeze(1)
src, tgt = drr.affine_inverse(src).contiguous(), drr.affine_inverse(tgt).contiguous()
res = []
for fn in (lambda s, t: drr.renderer(drr.density, s, t, raylen), lambda s, t: oracle.siddon_render(drr.density, s, t, raylen)):
    s, t = src.clone().requires_grad_(), tgt.clone().requires_grad_()
    img = fn(s, t)
    img.sum().backward()
    res.append((img.detach(), s.grad, t.grad))
print("img diff", (res[0][0] - res[1][0]).abs().max().item())
print("gsrc", res[0][1], res[1][1])
dt = (res[0][2] - res[1][2]).norm(dim=-1)
print("gtgt max diff per pose", dt.max(dim=1))
b = 2
worst = dt[b].topk(5).indices
for n in worst.tolist():
    print("ray", n, "ours", res[0][2][b, n].tolist(), "oracle", res[1][2][b, n].tolist())
    a = oracle.siddon_alphas(src[b:b+1], tgt[b:b+1, n:n+1], tuple(drr.density.shape), 0.5, 1e-8)[0, 0]
    a = a[~a.isnan()]
    print("   n_alpha", a.numel(), "min gap", torch.diff(a).min().item(), "src", src[b,0].tolist(), "tgt", tgt[b,n].tolist())
    idx, seg = oracle.siddon_segments(tuple(drr.density.shape), src[b:b+1], tgt[b:b+1, n:n+1])
    print("   idx", idx[0,0,:a.numel()-1].tolist())
