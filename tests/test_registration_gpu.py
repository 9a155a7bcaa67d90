"""BASELINE config 3 in miniature: GradNCC + mNCC + Adam pose refinement against a DRR of the same volume."""

import pytest
import torch

import oracle
import xvr_b200
from tests._scene import make_drr, oracle_render, rel_l2
from xvr_b200.registrar import Registrar

pytestmark = pytest.mark.gpu


def _problem(cuda, n=96, h=64):
    drr = make_drr(n, h)
    rot0 = torch.tensor([[0.20, -0.10, 0.05]], device=cuda)
    xyz0 = torch.tensor([[5.0, 800.0, -10.0]], device=cuda)
    with torch.no_grad():
        gt = drr(xvr_b200.convert(rot0, xyz0, parameterization="euler_angles", convention="ZXY"))
    d = torch.tensor([[0.06, -0.05, 0.04]], device=cuda)  # ~3 degrees
    t = torch.tensor([[8.0, 12.0, -6.0]], device=cuda)
    init = xvr_b200.convert(rot0 + d, xyz0 + t, parameterization="euler_angles", convention="ZXY")
    return drr, gt, init, rot0, xyz0


def test_registration_recovers_the_pose(cuda):
    drr, gt, init, rot0, xyz0 = _problem(cuda)
    reg = Registrar(drr, scales="1", n_itrs="300", patience=10, max_n_plateaus=3, use_cuda_graph=True)
    pose, info = reg.run(gt, init)
    rot, xyz = pose.convert("euler_angles", "ZXY")
    assert info["nccs"][-1] > 0.99 and info["nccs"][-1] > info["nccs"][0]
    assert (rot - rot0).abs().max().item() < 0.01 and (xyz - xyz0).abs().max().item() < 2.0
    n = info["n_itrs"][0]
    assert len(info["params"]) == n + 1 and len(info["nccs"]) == n + 1 and len(info["alphas"]) == n + 1
    assert len(info["times"]) == n + 1


def test_graph_replay_equals_eager_loop(cuda):
    res = []
    for graph in (True, False):
        drr, gt, init, *_ = _problem(cuda)
        pose, info = Registrar(drr, scales="1", n_itrs="40", use_cuda_graph=graph, poll_every=7).run(gt, init)
        res.append((pose.matrix.clone(), info))
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1]["nccs"] == res[1][1]["nccs"] and res[0][1]["alphas"] == res[1][1]["alphas"]


def test_trajectory_matches_the_reference_loop_on_the_oracle(cuda):
    """The loop of registrar/base.py:221-278 (torch Adam + ReduceLROnPlateau, .item() every iteration) run on the
    oracle renderer and metrics, against our graph-captured loop."""
    drr, gt, init, *_ = _problem(cuda, n=64, h=48)
    n_itr = 25
    pose, info = Registrar(drr, scales="1", n_itrs=str(n_itr), patience=4).run(gt, init)

    rot, xyz = init.convert("euler_angles", "ZXY")
    rot, xyz = torch.nn.Parameter(rot.clone()), torch.nn.Parameter(xyz.clone())
    opt = torch.optim.Adam([{"params": [rot], "lr": 1e-2}, {"params": [xyz], "lr": 1.0}], maximize=True)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, patience=4, threshold=1e-4, mode="max")
    img = oracle.xray_transforms(gt, 48)
    nccs = []
    for _ in range(n_itr):
        opt.zero_grad()
        pred = oracle.xray_transforms(oracle_render(drr, rot, xyz), 48)
        loss = (0.5 * oracle.multiscale_ncc(img, pred, (None, 9), (0.5, 0.5)) + 0.5 * oracle.gradient_ncc(img, pred, 11, 0.0)).sum()
        loss.backward()
        opt.step()
        sched.step(loss.detach())
        nccs.append(loss.item())
    ours = torch.tensor(info["nccs"][:n_itr])
    assert (ours - torch.tensor(nccs)).abs().max().item() < 2e-3
    final = torch.tensor(info["params"][-1])
    assert (final[:3] - rot.detach().cpu().flatten()).abs().max().item() < 2e-3
    assert (final[3:] - xyz.detach().cpu().flatten()).abs().max().item() < 0.2
    _ = rel_l2


@pytest.mark.parametrize("convention", ["ZXY", "XYZ", "YZX"])
def test_one_launch_euler_camera_matches_the_tensor_chain(cuda, convention, monkeypatch):
    """DRR.forward(rot, xyz, parameterization="euler_angles") builds the camera matrices in one launch
    (csrc/regstep.cu); image and pose gradients must equal the convert -> compose -> affine-inverse tensor chain."""
    drr = make_drr(48, 32)
    rot = torch.tensor([[0.30, -0.20, 0.10], [-0.15, 0.25, 0.05]], device=cuda)
    xyz = torch.tensor([[12.0, 790.0, -8.0], [-20.0, 860.0, 15.0]], device=cuda)
    w = torch.rand(2, 1, 32, 32, generator=torch.Generator().manual_seed(2)).to(cuda)
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("XVR_B200_FUSED_POSE", fused)
        r, x = rot.clone().requires_grad_(), xyz.clone().requires_grad_()
        img = drr(r, x, parameterization="euler_angles", convention=convention)
        (img * w).sum().backward()
        out[fused] = (img.detach(), r.grad, x.grad)
    assert rel_l2(out["1"][0], out["0"][0]) < 1e-5
    assert rel_l2(out["1"][1], out["0"][1]) < 2e-3 and rel_l2(out["1"][2], out["0"][2]) < 2e-3
    # degrees=True goes through the same kernel
    monkeypatch.setenv("XVR_B200_FUSED_POSE", "1")
    img_deg = drr(torch.rad2deg(rot), xyz, parameterization="euler_angles", convention=convention, degrees=True)
    assert rel_l2(img_deg, out["1"][0]) < 1e-5


def test_fused_update_matches_tensor_op_update(cuda):
    """xvr_reg_update (Adam + plateau scheduler + stopping rule + log row in one launch) against the tensor-op
    implementations that tests/test_cpu_registrar.py pins to torch.optim, step by step on the SAME scripted
    gradients and similarities (plateaus, a recovery, then the stop): parameters, moments, scheduler state and the
    trajectory log must agree after every update."""
    import ctypes

    from xvr_b200._lib import call, ptr, stream
    from xvr_b200.registrar import PlateauScheduler, adam_maximize_

    g = torch.Generator().manual_seed(5)
    n_steps, patience, max_plateaus = 40, 2, 4
    losses = torch.cat([torch.linspace(0.5, 0.8, 8), torch.full((9,), 0.8), torch.tensor([0.9]), torch.full((22,), 0.9)])
    grads = [(torch.randn(1, 3, generator=g).to(cuda), torch.randn(1, 3, generator=g).to(cuda) * 5) for _ in range(n_steps)]

    def fresh():
        packed = torch.zeros(8, device=cuda, dtype=torch.float64)
        packed[6], packed[7] = 1e-2, 1.0
        sched = PlateauScheduler([packed[6], packed[7]], patience=patience, max_n_plateaus=max_plateaus, storage=packed[1:6])
        rot = torch.tensor([[0.1, -0.2, 0.3]], device=cuda)
        xyz = torch.tensor([[5.0, 800.0, -10.0]], device=cuda)
        m = [torch.zeros_like(rot), torch.zeros_like(xyz)]
        v = [torch.zeros_like(rot), torch.zeros_like(xyz)]
        rows = torch.zeros(n_steps, 9, device=cuda)
        count = torch.zeros((), device=cuda)
        return packed, sched, rot, xyz, m, v, rows, count

    A = fresh()
    B = fresh()
    hyper = (ctypes.c_double * 9)(0.9, 0.999, 1e-8, 0.1, patience, 1e-4, 0.0, 1e-8, max_plateaus)
    for i in range(n_steps):
        loss = losses[i].to(cuda)
        # fused
        packed, sched, rot, xyz, m, v, rows, count = A
        call("xvr_reg_update", ptr(rot), ptr(xyz), ptr(grads[i][0]), ptr(grads[i][1]), 3, ptr(m[0]), ptr(v[0]), ptr(m[1]),
             ptr(v[1]), ptr(packed), ptr(loss), ptr(rows), ptr(count), n_steps, hyper, stream())
        # tensor ops, in the order of Registrar._iteration
        packed, sched, rot, xyz, m, v, rows, count = B
        state = {"step": packed[0], "exp_avg": m, "exp_avg_sq": v}
        adam_maximize_([rot, xyz], list(grads[i]), state, sched.lrs, active=sched.active)
        was_active = sched.active.clone()
        sched.step(loss)
        if was_active.item() > 0:
            rows[int(count.item())] = torch.cat([loss.reshape(1), rot.reshape(-1), xyz.reshape(-1), torch.stack(sched.lrs).float()])
            count += 1
        for name, x, y in (("state", A[0], B[0]), ("rot", A[2], B[2]), ("xyz", A[3], B[3]), ("m_rot", A[4][0], B[4][0]),
                           ("v_xyz", A[5][1], B[5][1]), ("rows", A[6], B[6]), ("count", A[7], B[7])):
            assert torch.allclose(x.double(), y.double(), rtol=2e-6, atol=1e-9), (i, name, x, y)
    assert A[0][5].item() == 0.0 and A[7].item() < n_steps  # the stopping rule fired and froze the state
    assert A[0][4].item() == max_plateaus


def test_registrar_follows_the_genuine_reference_loop(cuda):
    """tests/golden/reference_loop_v1.pt: xvr's own run_test_time_optimization (unmodified) on the oracle renderer.
    Our graph-captured loop on the CUDA renderer must see the same similarities, drop the learning rates at the same
    iterations, stop after the same iteration and end at the same pose."""
    import os

    from tests.golden.make_golden import scene
    from xvr_b200.data import read

    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_loop_v1.pt"), weights_only=False)
    sc, h = gold["scene"], gold["hyper"]
    hu, _, affine = scene(sc["n"])
    drr = xvr_b200.DRR(read(hu, affine=affine, center_volume=False), sc["sdd"], sc["height"], sc["delx"],
                       renderer="trilinear", reverse_x_axis=False).to(cuda)
    rot0, xyz0 = gold["rot0"].to(cuda), gold["xyz0"].to(cuda)
    with torch.no_grad():
        gt = drr(xvr_b200.convert(rot0, xyz0, parameterization="euler_angles", convention="ZXY"))
    assert rel_l2(gt.cpu(), gold["gt"]) < 1e-4
    init = xvr_b200.convert(rot0 + gold["drot"].to(cuda), xyz0 + gold["dxyz"].to(cuda), parameterization="euler_angles",
                            convention="ZXY")
    reg = Registrar(drr, scales="1", n_itrs=str(h["n_itrs"][0]), lr_rot=h["lr_rot"], lr_xyz=h["lr_xyz"],
                    patience=h["patience"], threshold=h["threshold"], max_n_plateaus=h["max_n_plateaus"])
    _, info = reg.run(gt, init)
    n = len(gold["nccs"]) - 1
    assert info["n_itrs"] == [n]
    assert (torch.tensor(info["nccs"], dtype=torch.float64) - gold["nccs"]).abs().max().item() < 2e-3
    assert torch.allclose(torch.tensor(info["alphas"], dtype=torch.float64), gold["alphas"], rtol=1e-6)
    ours = torch.tensor(info["params"], dtype=torch.float64)
    assert (ours[:, :3] - gold["params"][:, :3]).abs().max().item() < 2e-3
    assert (ours[:, 3:] - gold["params"][:, 3:]).abs().max().item() < 0.2
