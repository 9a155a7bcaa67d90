"""The plateau/stop logic of our registration loop against a run of the GENUINE xvr loop
(tests/golden/make_reference_loop_golden.py: registrar/base.py:198-292 unmodified, on the oracle renderer)."""

import os

import torch

from xvr_b200.registrar import PlateauScheduler

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_loop_v1.pt"), weights_only=False)


def test_plateau_scheduler_replays_the_reference_run():
    """Fed the similarities the reference loop saw, our scheduler must produce the learning rates it logged after
    every iteration and stop after the same iteration (the reference counts its first iteration as a plateau)."""
    h = GOLD["hyper"]
    nccs, alphas = GOLD["nccs"], GOLD["alphas"]
    n = len(nccs) - 1  # the last entry is the final evaluation after the loop
    assert alphas.shape == (n + 1, 2) and n < h["n_itrs"][0]  # the stop rule ended the run, not the iteration cap
    lrs = [torch.tensor(h["lr_rot"], dtype=torch.float64), torch.tensor(h["lr_xyz"], dtype=torch.float64)]
    sched = PlateauScheduler(lrs, factor=0.1, patience=h["patience"], threshold=h["threshold"],
                             max_n_plateaus=h["max_n_plateaus"])
    for i in range(n):
        assert sched.active > 0, f"stopped before iteration {i}, the reference ran {n}"
        sched.step(nccs[i].to(torch.float32))  # the loop hands the scheduler an fp32 similarity
        assert abs(float(lrs[0]) - float(alphas[i + 1, 0])) < 1e-15
        assert abs(float(lrs[1]) - float(alphas[i + 1, 1])) < 1e-15
    assert sched.active == 0 and float(sched.n_plateaus) == h["max_n_plateaus"]


def test_parse_scales_matches_the_reference():
    from xvr_b200.registrar import parse_scales

    for (scales, crop, height), ref in GOLD["parse_scales"].items():
        assert parse_scales(scales, crop, height) == ref
