"""CPU: the LR schedule mirror (hsimae_b200.optim.CosineLRScheduler).  timm is not installed, so this pins the
restated formula's own properties (unpinned against timm), not timm's output."""
import math

import torch


def test_cosine_schedule_shape():
    from hsimae_b200.optim import CosineLRScheduler
    p = torch.zeros(1, requires_grad=True)
    opt = torch.optim.SGD([{"params": [p], "lr": 5e-3}], lr=5e-3)
    iters = 200
    s = CosineLRScheduler(opt, t_initial=iters, lr_min=1e-6, warmup_t=int(math.ceil(iters * 0.05)))   # Model_Pretraining.py:88
    assert opt.param_groups[0]["lr"] == 0.0                       # starts at warmup_lr_init
    lrs = []
    for t in range(iters):
        s.step(t)
        lrs.append(opt.param_groups[0]["lr"])
    assert lrs[0] == 0.0 and abs(lrs[5] - 5e-3 * 5 / 10) < 1e-12   # linear warm-up over 10 steps
    assert abs(lrs[10] - (1e-6 + 0.5 * (5e-3 - 1e-6) * (1 + math.cos(math.pi * 10 / iters)))) < 1e-12
    assert all(x >= y for x, y in zip(lrs[10:], lrs[11:]))         # monotone decay after the warm-up
    assert 1e-6 <= lrs[-1] < 1e-5
    s.step(iters + 5)
    assert opt.param_groups[0]["lr"] == 1e-6
