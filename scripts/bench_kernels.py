"""Per-kernel timings and roofline fractions for every CUDA entry point on the hot path.

    python scripts/bench_kernels.py [--quick]     -> JSON lines + a markdown table on stdout
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, xvr_b200
from xvr_b200 import _lib, metrics
from xvr_b200.data import read, synthetic_ct

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="trilinear,siddon,ncc")
args = ap.parse_args()
dev = torch.device("cuda")
PEAK, _ = bench.peaks()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []


def report(name, ms, alg_bytes, units, unit_name, note=""):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    rec = {"kernel": name, "ms": ms, "algorithmic_GB": alg_bytes / 1e9, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / PEAK,
           "throughput": units / (ms * 1e-3), "unit": unit_name, "note": note}
    rows.append(rec)
    print(json.dumps(rec), flush=True)


def rays(drr, rot, xyz):
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    src, tgt = drr.detector(pose, None)
    raylen = (tgt - src).norm(dim=-1).unsqueeze(1).contiguous()
    return drr.affine_inverse(src).contiguous(), drr.affine_inverse(tgt).contiguous(), raylen


ONLY = args.only.split(",")


def section_trilinear():
    B = 16 if args.quick else 116
    drr = bench.build_scene(dev, 512, 256)
    rot, xyz = (t.to(dev) for t in bench.pose_batch(B, 0))
    A_fwd = bench.algorithmic_bytes_fwd(256, 256, 500)
    pose = xvr_b200.convert(rot, xyz, parameterization="euler_angles", convention="ZXY")
    with torch.no_grad():
        report("trilinear fused fwd (no grad)", timeit(lambda: drr(pose)), B * A_fwd, B, "DRR/s", "512^3, 256^2, n=500")
    r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
    def fwd_jac():
        return drr(xvr_b200.convert(r, x, parameterization="euler_angles", convention="ZXY"))
    report("trilinear fused fwd + Jacobian", timeit(fwd_jac), B * A_fwd, B, "DRR/s", "one march for fwd+bwd(pose)")
    src, tgt, raylen = rays(drr, rot, xyz)
    with torch.no_grad():
        report("trilinear rays fwd (materialised rays)", timeit(lambda: drr.renderer(drr.density, src, tgt, raylen)), B * A_fwd, B, "DRR/s")
    s_, t_ = src.clone().requires_grad_(), tgt.clone().requires_grad_()
    gout = torch.rand(B, 1, 65536, device=dev)
    def rays_fb():
        drr.renderer(drr.density, s_, t_, raylen).backward(gout)
    report("trilinear rays fwd+bwd (Jacobian path)", timeit(rays_fb), 2 * B * A_fwd, B, "DRR/s", "SURVEY two-pass bytes")
    # label channels
    hu, lab, aff = synthetic_ct(512, with_labels=True, device=dev)
    labels = lab
    C = int(labels.max()) + 1
    with torch.no_grad():
        report(f"trilinear rays fwd with {C} label channels", timeit(lambda: drr.renderer(drr.density, src, tgt, raylen, mask=labels)),
               B * 256 * 256 * (500 * 33 + 4 * C), B, "DRR/s", "+1 B/sample label gather")
    gl = torch.rand(B, C, 65536, device=dev)
    def lab_fb():
        drr.renderer(drr.density, s_, t_, raylen, mask=labels).backward(gl)
    report("trilinear rays fwd+bwd with label channels (recompute bwd)", timeit(lab_fb, reps=3), 2 * B * 256 * 256 * (500 * 33 + 4 * C), B, "DRR/s")
    del hu, lab, labels
    # volume gradient
    Bv = 8 if args.quick else 116
    vol = drr.density.detach().clone().requires_grad_()
    keep = drr.density
    drr.density = vol
    pv = xvr_b200.convert(rot[:Bv], xyz[:Bv], parameterization="euler_angles", convention="ZXY")
    img = drr(pv)
    def volgrad():
        vol.grad = None
        img.backward(gout[:Bv].view_as(img), retain_graph=True)
    report("trilinear dL/dvolume (brick-local scatter, atomics-free)", timeit(volgrad, reps=2, warm=1), Bv * 256 * 256 * 500 * 64, Bv, "DRR/s", "A_bwd_vol = 64 B/sample")
    drr.density = keep
    del vol, img, drr


def section_siddon():
    # ---------------------------------------------------------------- Siddon, config 5 (768^3, 512^2)
    Bs = 4 if args.quick else 32
    hu, _, aff = synthetic_ct(768, device=dev)
    sub = read(hu, affine=aff)
    del hu
    sdrr = xvr_b200.DRR(sub, bench.SDD, 512, bench.DELX / 2, renderer="siddon", reverse_x_axis=False).to(dev)
    rot, xyz = (t.to(dev) for t in bench.pose_batch(Bs, 1))
    src, tgt, raylen = rays(sdrr, rot, xyz)
    N = 512 * 512
    cnt = torch.zeros(Bs, N, dtype=torch.int32, device=dev)
    idx = torch.zeros(Bs, N, 1, dtype=torch.int32, device=dev)
    seg = torch.zeros(Bs, N, 1, device=dev)
    _lib.call("xvr_siddon_trace", _lib.ptr(sdrr.density), None, *sdrr.density.shape, _lib.ptr(src), _lib.ptr(tgt), Bs, N, 0.5, 1e-8, 1,
              _lib.ptr(idx), _lib.ptr(seg), _lib.ptr(cnt), _lib.opts_word(), _lib.stream())
    nseg = cnt.sum().item()
    print(json.dumps({"siddon_mean_segments_per_ray": nseg / (Bs * N)}))
    A_sid = nseg * 4 + Bs * N * 4
    with torch.no_grad():
        report("siddon rays fwd", timeit(lambda: sdrr.renderer(sdrr.density, src, tgt, raylen), reps=3), A_sid, Bs, "DRR/s", "768^3, 512^2")
    s_, t_ = src.clone().requires_grad_(), tgt.clone().requires_grad_()
    gs = torch.rand(Bs, 1, N, device=dev)
    def sid_fb():
        sdrr.renderer(sdrr.density, s_, t_, raylen).backward(gs)
    report("siddon rays fwd+bwd (Jacobian path)", timeit(sid_fb, reps=3), 2 * A_sid, Bs, "DRR/s", "SURVEY two-pass bytes")



def section_ncc():
    # ---------------------------------------------------------------- similarity, config 2 sizes
    Bn = 116
    x1, x2 = torch.randn(Bn, 1, 256, 256, device=dev), torch.randn(Bn, 1, 256, 256, device=dev).requires_grad_()
    mncc = metrics.MultiscaleNormalizedCrossCorrelation2d([None, 9], [0.5, 0.5])
    gncc = metrics.GradientNormalizedCrossCorrelation2d(11, 0.0).to(dev)
    A_ncc = Bn * 256 * 256 * 8
    with torch.no_grad():
        report("mNCC([None,9]) fwd", timeit(lambda: mncc(x1, x2)), A_ncc, Bn, "img/s")
        report("GradNCC(11) fwd", timeit(lambda: gncc(x1, x2)), A_ncc, Bn, "img/s")
    def m_fb():
        x2.grad = None
        mncc(x1, x2).sum().backward()
    def g_fb():
        x2.grad = None
        gncc(x1, x2).sum().backward()
    report("mNCC([None,9]) fwd+bwd", timeit(m_fb), 2 * A_ncc, Bn, "img/s")
    report("GradNCC(11) fwd+bwd", timeit(g_fb), 2 * A_ncc, Bn, "img/s")


if "trilinear" in ONLY:
    section_trilinear()
    torch.cuda.empty_cache()
if "siddon" in ONLY:
    section_siddon()
    torch.cuda.empty_cache()
if "ncc" in ONLY:
    section_ncc()

print("\n| kernel | ms | algorithmic GB | achieved GB/s | frac of 6450 GB/s | throughput | note |")
print("|---|---|---|---|---|---|---|")
for r_ in rows:
    print(f"| {r_['kernel']} | {r_['ms']:.3f} | {r_['algorithmic_GB']:.2f} | {r_['achieved_GBps']:.0f} | {r_['frac_of_hbm_peak']:.3f} | {r_['throughput']:.1f} {r_['unit']} | {r_['note']} |")
