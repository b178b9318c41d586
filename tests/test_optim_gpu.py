"""GPU: fused AdamW (hsimae_adamw_step via hsimae_b200.optim.FusedAdamW) against torch.optim.AdamW configured the way the
reference drivers do (/root/reference/Model_Pretraining.py:80-86).  fp32 element-wise arithmetic in torch's operation
order: the bar is agreement to a few ulp (the tolerance below), in practice bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    shapes = [(256, 256), (684, 256), (256,), (1,), (3, 5, 7), (4099,), (72, 256), (17, 1024)]
    return [torch.randn(*s, device=DEV, generator=g).requires_grad_(True) for s in shapes]


def _groups(ps):
    return [{"params": ps[0::2], "weight_decay": 5e-2}, {"params": ps[1::2], "weight_decay": 0.0}]


@pytest.mark.parametrize("lr,betas", [(5e-3, (0.9, 0.95)), (1e-3, (0.9, 0.999))])
def test_matches_torch_adamw_over_steps(lr, betas):
    from hsimae_b200.optim import FusedAdamW
    pa, pb = _params(0), _params(0)
    ref = torch.optim.AdamW(_groups(pa), lr=lr, weight_decay=5e-2, betas=betas, foreach=True)
    ours = FusedAdamW(_groups(pb), lr=lr, weight_decay=5e-2, betas=betas)
    g = torch.Generator(device=DEV).manual_seed(1)
    worst = 0.0
    for step in range(12):
        if step == 6:                       # the schedulers change the rate between steps
            for opt in (ref, ours):
                for grp in opt.param_groups:
                    grp["lr"] = lr * 0.37
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g) * (0.1 if step % 3 else 3.0)
            a.grad = gr.clone(); b.grad = gr.clone()
        if step == 9:                       # parameters without a gradient are left alone (frozen pos_embed / mask_token)
            pa[3].grad = None; pb[3].grad = None
        ref.step(); ours.step()
        for a, b in zip(pa, pb):
            worst = max(worst, float(((a - b).abs() / (a.abs() + 1e-6)).max()))
            assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (step, a.shape)
    for a, b in zip(pa, pb):
        sa, sb = ref.state[a], ours.state[b]
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=2e-6, atol=1e-9)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    print("worst relative difference to torch.optim.AdamW: %.3g" % worst)


def test_training_loop_with_fused_optimizer_and_schedule():
    """Model_Pretraining.py:80-104 with the opt-in optimiser / scheduler: same loss curve as torch's AdamW."""
    import random
    import Models
    from hsimae_b200.optim import FusedAdamW, CosineLRScheduler

    def run(fused):
        torch.manual_seed(0); random.seed(0)
        model = Models.HSIMAE(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=128, depth=12, num_heads=8, s_depth=9,
                              decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True).to(DEV)
        no_decay = ["bias", "norm"]
        groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
                  {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
        opt = (FusedAdamW if fused else torch.optim.AdamW)(groups, lr=1e-3, weight_decay=5e-2, betas=(0.9, 0.95))
        sched = CosineLRScheduler(opt, t_initial=16, lr_min=1e-6, warmup_t=2)
        x = torch.randn(64, 1, 32, 9, 9, device=DEV)
        losses = []
        for it in range(16):
            loss, _, _ = model(x, mask_ratio=0.5)
            opt.zero_grad()
            loss.backward()
            opt.step()
            sched.step(it)
            losses.append(loss.item())
        return losses

    a, b = run(False), run(True)
    print("torch AdamW :", ["%.5f" % v for v in a])
    print("fused AdamW :", ["%.5f" % v for v in b])
    assert max(abs(x - y) for x, y in zip(a, b)) < 2e-3, (a, b)
    assert min(b[3:]) < b[0]


def test_errors():
    from hsimae_b200.optim import FusedAdamW
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()
