"""GPU: the CUDA-graph fine-tuning step (hsimae_b200/graph.py) trains exactly like the eager loop body it replaces
(Model_Finetuning.py:147-166) from the same seeds: same visible shapes (Python RNG), same mask noise and
stochastic-depth draws (torch CUDA generator, advanced by every replay)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _train(graphed: bool, steps: int = 6):
    import Models
    from hsimae_b200.graph import GraphedFinetuneStep
    torch.manual_seed(5); random.seed(5)
    model = Models.DualViT(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=128, depth=12,
                           num_heads=8, s_depth=9, drop_path=0.2, decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=8,
                           norm_pix_loss=True, trunc_init=True).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-2)
    crit = torch.nn.CrossEntropyLoss(ignore_index=0)
    g = torch.Generator(device="cpu").manual_seed(1)
    batches = [(torch.randn(32, 1, 32, 9, 9, generator=g).cuda(), torch.randn(71, 1, 32, 9, 9, generator=g).cuda(),
                torch.randint(0, 17, (32,), generator=g).cuda()) for _ in range(steps)]
    stepper = GraphedFinetuneStep(model, opt, crit, lamda=10.0, mask_ratio=0.8, warmup=0) if graphed else None
    torch.manual_seed(77); random.seed(77)
    losses, shapes = [], []
    for x, xu, y in batches:
        if graphed:
            loss, logits = stepper(x, xu, y)
        else:
            loss_rec, _, _, logits = model(x, xu, mask_ratio=0.8)
            loss = 10.0 * loss_rec + crit(logits, y)
            opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
        shapes.append((int(model.len_t), int(model.len_l)) if not graphed else None)
    return losses, {k: v.detach().clone() for k, v in model.state_dict().items()}, stepper


def test_graphed_finetune_step_matches_eager():
    eager_losses, eager_sd, _ = _train(False)
    graph_losses, graph_sd, stepper = _train(True)
    assert len(stepper.graphs) >= 1
    # warmup=0: capture consumes no extra generator draws, so the two runs see the same noise; reductions are atomics
    for a, b in zip(eager_losses, graph_losses):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (eager_losses, graph_losses)
    # attn.k.bias has a mathematically zero gradient (softmax is shift invariant): AdamW turns its rounding noise into +-lr steps
    worst = max(float((eager_sd[k] - graph_sd[k]).norm() / (eager_sd[k].norm() + 1e-6)) for k in eager_sd if not k.endswith("attn.k.bias"))
    assert worst < 2e-2, worst
