"""CPU suite, part 4: the device-side plateau scheduler / Adam used by the graph-captured registration loop,
against torch.optim.Adam + ReduceLROnPlateau and the reference loop's stopping rule (registrar/base.py:221-278)."""

import torch

from xvr_b200.registrar import PlateauScheduler, adam_maximize_, parse_scales


def test_parse_scales():
    assert parse_scales("8", 0, 256) == [1 / 8]
    assert parse_scales("8,4,2", 0, 256) == [1 / 8, 2.0, 2.0]
    assert parse_scales("1", 0, 256) == [1.0]
    assert abs(parse_scales("8", 100, 1436)[0] - 1 / (8 * 1436 / 1536)) < 1e-12


def test_plateau_scheduler_matches_torch_and_reference_stop_rule():
    torch.manual_seed(0)
    for trial in range(6):
        seq = (torch.rand(90).cumsum(0) * 0.01 + torch.randn(90) * 0.02).clamp(max=0.45 + 0.1 * trial)
        p, q = torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([{"params": [p], "lr": 1e-2}, {"params": [q], "lr": 1.0}], maximize=True)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, patience=3, threshold=1e-4, mode="max")
        lrs = [torch.tensor(1e-2, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64)]
        mine = PlateauScheduler(lrs, patience=3, max_n_plateaus=3)
        n_plateaus, current, stop_ref, stop_mine = 0, float("inf"), None, None
        for i, v in enumerate(seq):
            ref.step(v)
            lr = ref.get_last_lr()
            if stop_ref is None:
                if lr[0] < current:
                    current, n_plateaus = lr[0], n_plateaus + 1
                if n_plateaus == 3:
                    stop_ref = i
            if mine.active > 0:
                mine.step(v)
                assert abs(float(lrs[0]) - lr[0]) < 1e-12 and abs(float(lrs[1]) - lr[1]) < 1e-12
                if mine.active == 0:
                    stop_mine = i
            else:  # frozen after the stop
                before = [t.clone() for t in mine.state()]
                mine.step(v)
                assert all(torch.equal(a, b) for a, b in zip(before, mine.state()))
        assert stop_ref == stop_mine and stop_ref is not None


def test_adam_maximize_matches_torch():
    p = torch.nn.Parameter(torch.tensor([[0.3, -0.2, 0.1]]))
    opt = torch.optim.Adam([p], lr=0.05, maximize=True)
    q = p.detach().clone()
    st = {"step": torch.zeros((), dtype=torch.float64), "exp_avg": [torch.zeros_like(q)], "exp_avg_sq": [torch.zeros_like(q)]}
    on = torch.ones((), dtype=torch.float64)
    for i in range(40):
        g = torch.sin(torch.arange(3.0) + i)[None]
        p.grad = g.clone()
        opt.step()
        adam_maximize_([q], [g], st, [torch.tensor(0.05, dtype=torch.float64)], active=on)
    assert (p.detach() - q).abs().max().item() < 1e-6
    frozen = q.clone()
    adam_maximize_([q], [g], st, [torch.tensor(0.05, dtype=torch.float64)], active=torch.zeros((), dtype=torch.float64))
    assert torch.equal(q, frozen) and float(st["step"]) == 40.0
