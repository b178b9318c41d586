"""GPU: whole-model parity of the CUDA path (through Models.py -> C ABI) against the CPU fp32 oracle on the
same weights, inputs and noise, and against the fixtures the reference produced.

Tolerances (bf16 tensor-core GEMM operands, fp32 accumulation / residual stream / statistics):
  mask indices ............ bit-exact
  loss .................... |d| <= 3e-3 * |loss|
  pred / logits / latent .. ||d|| / ||ref|| <= 2e-2
  parameter gradients ..... ||d|| / ||ref|| <= 6e-2 per tensor (LayerNorm/bias vectors: 1e-1)
"""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import TINY, grads_from_npz, rel_err, state_from_npz, tiny_geometry
from oracle import hsimae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_LOSS, TOL_ACT, TOL_GRAD, TOL_GRAD_VEC = 3e-3, 2e-2, 6e-2, 1e-1


def _check_grads(model, ref_grads, skip=()):
    bad = []
    named = dict(model.named_parameters())
    for k, g in ref_grads.items():
        if k in skip:
            continue
        assert named[k].grad is not None, f"no gradient for {k}"
        if k.endswith("attn.k.bias"):
            # softmax is invariant to a shift of every key score, so d/d(k.bias) is exactly zero in real arithmetic;
            # both sides only hold rounding noise -- compare against the scale of the sibling q.bias gradient instead
            scale = float(ref_grads[k.replace("attn.k.bias", "attn.q.bias")].norm())
            assert float(named[k].grad.norm()) <= 0.05 * scale + 1e-7, k
            continue
        e = rel_err(named[k].grad, g)
        tol = TOL_GRAD_VEC if g.dim() == 1 else TOL_GRAD
        if not e < tol:
            bad.append((k, e))
    for k, p in named.items():
        if k not in ref_grads:
            assert p.grad is None, f"unexpected gradient for {k}"
    assert not bad, "gradient mismatch: " + ", ".join(f"{k}:{e:.3g}" for k, e in bad[:12])


def _drops_as_dict(lst, s_depth):
    out = {}
    if lst is None:
        return out
    n_f = (len(lst) - 4 * s_depth) // 2
    for stack, base, cnt in ((1, 0, s_depth), (2, 2 * s_depth, s_depth), (0, 4 * s_depth, n_f)):
        for i in range(cnt):
            for j in (0, 1):
                t = lst[base + 2 * i + j]
                if t is not None:
                    out[(stack, i, j + 1)] = t.detach().cpu()
    return out


def test_pretrain_matches_reference_fixture(golden):
    import Models as M
    z = golden("tiny_pretrain.npz")
    model = M.HSIMAE(**TINY)
    model.load_state_dict(state_from_npz(z))
    model = model.to(DEV)
    x = torch.from_numpy(z["x"]).to(DEV)
    # find a seed whose python draw gives the fixture's visible shape, then check against the oracle on our own noise
    g = tiny_geometry()
    sd = state_from_npz(z)
    for seed in range(20):
        random.seed(seed); torch.manual_seed(seed)
        loss, pred, mask = model(x, mask_ratio=0.5)
        aux = model._last
        if (aux["lt"], aux["ll"]) == (int(z["len_t"]), int(z["len_l"])):
            break
    model.zero_grad()
    loss.backward()
    out, grads = O.pretrain_step_grads(sd, x.cpu(), g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"])
    assert torch.equal(aux["ids_keep"].cpu(), out["ids_keep"]) and torch.equal(aux["ids_restore"].cpu(), out["ids_restore"])
    assert torch.equal(mask.cpu(), out["mask_img"])
    assert abs(loss.item() - out["loss"].item()) <= TOL_LOSS * abs(out["loss"].item())
    assert rel_err(pred, out["pred_img"]) < TOL_ACT
    assert loss.dim() == 0 and pred.shape == x.shape and mask.shape == x.shape
    _check_grads(model, grads)
    # the reference's own numbers (its noise differs from ours, so compare distribution-level quantities)
    assert abs(loss.item() - float(z["loss"])) < 0.05


def test_pretrain_with_reference_noise(golden):
    """same noise as the reference run => identical mask, loss/pred/grads within tolerance of the REFERENCE outputs"""
    import Models as M
    from hsimae_b200 import _lib
    z = golden("tiny_pretrain.npz")
    model = M.HSIMAE(**TINY)
    model.load_state_dict(state_from_npz(z))
    model = model.to(DEV)
    x = torch.from_numpy(z["x"]).to(DEV)
    lt, ll = int(z["len_t"]), int(z["len_l"])
    import hsimae_b200.modules as mod
    orig_shape, orig_rand = mod.choose_visible_shape, torch.rand
    feed = [torch.from_numpy(z["noise_t"]).to(DEV), torch.from_numpy(z["noise_l"]).to(DEV)]
    mod.choose_visible_shape = lambda T, L, r: (lt, ll)
    torch.rand = lambda *a, **k: feed.pop(0)
    try:
        loss, pred, mask = model(x, mask_ratio=0.5)
    finally:
        mod.choose_visible_shape, torch.rand = orig_shape, orig_rand
    loss.backward()
    assert torch.equal(model._last["ids_keep"].cpu(), torch.from_numpy(z["ids_keep"]))
    assert torch.equal(model._last["ids_restore"].cpu(), torch.from_numpy(z["ids_restore"]))
    assert torch.equal(mask.cpu(), torch.from_numpy(z["mask"]))
    assert abs(loss.item() - float(z["loss"])) <= TOL_LOSS * float(z["loss"])
    assert rel_err(pred, torch.from_numpy(z["pred"])) < TOL_ACT
    _check_grads(model, grads_from_npz(z))


@pytest.mark.parametrize("dim,heads,n,ratio", [(128, 8, 48, 0.5), (256, 16, 40, 0.5), (128, 8, 21, 0.75), (256, 16, 19, 0.8),
                                               (144, 9, 10, 0.9)])
def test_pretrain_reference_configs(dim, heads, n, ratio):
    """Base / Large (Model_Pretraining.py:130) and the fine-tune default width, all mask ratios the reference uses"""
    import Models as M
    kw = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=dim, depth=12, num_heads=heads, s_depth=9,
              decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
    g = O.Geometry(embed_dim=dim, num_heads=heads)
    sd = O.make_state(g, seed=dim + n, randomize_affine=True)
    model = M.HSIMAE(**kw)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV)
    torch.manual_seed(n); random.seed(n)
    x = torch.randn(n, 1, 32, 9, 9, device=DEV)
    loss, pred, mask = model(x, mask_ratio=ratio)
    loss.backward()
    aux = model._last
    out, grads = O.pretrain_step_grads(sd, x.cpu(), g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"])
    assert torch.equal(aux["ids_keep"].cpu(), out["ids_keep"]) and torch.equal(mask.cpu(), out["mask_img"])
    assert abs(loss.item() - out["loss"].item()) <= TOL_LOSS * abs(out["loss"].item())
    assert rel_err(pred, out["pred_img"]) < TOL_ACT
    _check_grads(model, grads)


def test_forward_encoder_and_no_grad():
    import Models as M
    g = tiny_geometry()
    sd = O.make_state(g, seed=3, randomize_affine=True)
    model = M.HSIMAE(**TINY)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV).eval()
    x = torch.randn(12, 1, 32, 9, 9, device=DEV)
    random.seed(1); torch.manual_seed(1)
    with torch.no_grad():
        st = (torch.cuda.get_rng_state(), random.getstate())
        latent, mask, ids_restore, ids_keep = model.forward_encoder(x, 0.5)
        torch.cuda.set_rng_state(st[0]); random.setstate(st[1])
        loss, pred, m2 = model(x, mask_ratio=0.5)
    assert not loss.requires_grad
    lt, ll = int(model.len_t), int(model.len_l)
    ref = O.encode_masked(sd, x.cpu(), g, ids_keep.cpu(), lt, ll)
    assert rel_err(latent, ref) < TOL_ACT
    out = O.pretrain_forward(sd, x.cpu(), g, model._last["noise_t"].cpu(), model._last["noise_l"].cpu(), lt, ll)
    assert torch.equal(model._last["ids_keep"].cpu(), ids_keep.cpu())        # same RNG stream => same mask
    assert abs(loss.item() - out["loss"].item()) <= TOL_LOSS * out["loss"].item()


def test_dual_and_vit_match_reference_fixture(golden):
    import Models as M
    import hsimae_b200.modules as mod
    z = golden("tiny_dual.npz")
    kw = dict(TINY); kw.update(num_class=17, drop_path=0.0)
    model = M.DualViT(**kw)
    model.load_state_dict(state_from_npz(z))
    model = model.to(DEV).train()
    xl, xu = torch.from_numpy(z["xl"]).to(DEV), torch.from_numpy(z["xu"]).to(DEV)
    labels = torch.from_numpy(z["labels"]).to(DEV)
    lt, ll = int(z["len_t"]), int(z["len_l"])
    orig_shape, orig_rand = mod.choose_visible_shape, torch.rand
    feed = [torch.from_numpy(z["noise_t"]).to(DEV), torch.from_numpy(z["noise_l"]).to(DEV)]
    mod.choose_visible_shape = lambda T, L, r: (lt, ll)
    torch.rand = lambda *a, **k: feed.pop(0)
    try:
        loss_rec, pred_rec, mask, logits = model(xl, xu, mask_ratio=0.8)
    finally:
        mod.choose_visible_shape, torch.rand = orig_shape, orig_rand
    total = 10.0 * loss_rec + F.cross_entropy(logits, labels, ignore_index=0)
    total.backward()
    assert torch.equal(mask.cpu(), torch.from_numpy(z["mask"]))
    assert abs(loss_rec.item() - float(z["loss_rec"])) <= TOL_LOSS * float(z["loss_rec"])
    assert rel_err(logits, torch.from_numpy(z["logits"])) < TOL_ACT
    assert rel_err(pred_rec, torch.from_numpy(z["pred_rec"])) < TOL_ACT
    assert abs(total.item() - float(z["total"])) <= 5e-3 * float(z["total"])
    _check_grads(model, grads_from_npz(z))
    # eval path: class_pred only (Model_Finetuning.py:193)
    model.eval()
    with torch.no_grad():
        ev = model(xl, mask_ratio=0.8)
    assert ev.shape == (6, 17) and rel_err(ev, torch.from_numpy(z["logits_eval"])) < TOL_ACT
    # HSIViT loads the shared keys (Model_Finetuning.py:253-261)
    vkw = {k: v for k, v in kw.items() if not k.startswith("decoder") and k != "norm_pix_loss"}
    vit = M.HSIViT(**vkw)
    sdv = vit.state_dict()
    sdv.update({k: v for k, v in model.state_dict().items() if k in sdv})
    vit.load_state_dict(sdv)
    vit = vit.to(DEV).eval()
    with torch.no_grad():
        lv = vit(xl)
    assert rel_err(lv, torch.from_numpy(z["logits_vit"])) < TOL_ACT


def test_dual_drop_path_training_matches_oracle():
    """stochastic depth: per-(b,t) / per-(b,l) / per-b factors drawn on the device, fed to the oracle"""
    import Models as M
    g = tiny_geometry(17)
    sd = O.make_state(g, seed=9, head=True, randomize_affine=True)
    kw = dict(TINY); kw.update(num_class=17, drop_path=0.4)
    model = M.DualViT(**kw)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV).train()
    torch.manual_seed(5); random.seed(5)
    xl, xu = torch.randn(7, 1, 32, 9, 9, device=DEV), torch.randn(12, 1, 32, 9, 9, device=DEV)
    labels = torch.randint(0, 17, (7,), device=DEV)
    loss_rec, pred_rec, mask, logits = model(xl, xu, mask_ratio=0.8)
    (5.0 * loss_rec + F.cross_entropy(logits, labels, ignore_index=0)).backward()
    aux = model._last
    assert any(d is not None for d in aux["drops"]) and any(d is not None for d in aux["drops_full"])
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in sd.items()}
    out = O.dual_forward(leaves, xl.cpu(), xu.cpu(), g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"],
                         _drops_as_dict(aux["drops_full"], 2), _drops_as_dict(aux["drops"], 2))
    (5.0 * out["loss"] + F.cross_entropy(out["logits"], labels.cpu(), ignore_index=0)).backward()
    assert torch.equal(mask.cpu(), out["mask_img"])
    assert abs(loss_rec.item() - out["loss"].item()) <= TOL_LOSS * out["loss"].item()
    assert rel_err(logits, out["logits"]) < TOL_ACT
    _check_grads(model, {k: v.grad for k, v in leaves.items() if v.grad is not None})


def test_full_size_properties():
    """BASELINE config-2 shape (N=4096, Large): size-independent properties instead of the (slow) oracle"""
    import Models as M
    kw = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=256, depth=12, num_heads=16, s_depth=9,
              decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
    torch.manual_seed(42); random.seed(42)
    model = M.HSIMAE(**kw).to(DEV)
    x = torch.randn(4096, 1, 32, 9, 9, device=DEV)
    st = (torch.cuda.get_rng_state(), random.getstate())
    loss, pred, mask = model(x, mask_ratio=0.5)
    loss.backward()
    g_full = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    assert torch.isfinite(loss) and 0.9 < loss.item() < 1.2          # untrained model on N(0,1) data: ~1
    assert int(mask.sum().item()) == 4096 * 18 * 72
    ids = model._last["ids_restore"]
    assert torch.equal(torch.sort(ids, 1).values, torch.arange(36, device=DEV).expand(4096, -1))
    # linearity of the batch mean: loss/grads of the full batch == mean over two halves (same noise rows)
    aux = model._last
    import hsimae_b200.modules as mod
    orig_shape, orig_rand = mod.choose_visible_shape, torch.rand
    halves = []
    for sl in (slice(0, 2048), slice(2048, 4096)):
        feed = [aux["noise_t"][sl].contiguous(), aux["noise_l"][sl].contiguous()]
        mod.choose_visible_shape = lambda T, L, r: (aux["lt"], aux["ll"])
        torch.rand = lambda *a, **k: feed.pop(0)
        try:
            model.zero_grad()
            l, _, _ = model(x[sl], mask_ratio=0.5)
        finally:
            mod.choose_visible_shape, torch.rand = orig_shape, orig_rand
        l.backward()
        halves.append((l.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    assert abs(0.5 * (halves[0][0] + halves[1][0]) - loss.item()) < 1e-4
    for k in ("decoder_pred.weight", "blocks.2.mlp.w2.weight", "blocks_1.0.attn.q.weight", "patch_embed.proj.weight", "norm.bias"):
        assert rel_err(0.5 * (halves[0][1][k] + halves[1][1][k]), g_full[k]) < 2e-3, k


def test_training_loop_decreases_loss():
    """a few AdamW steps of the reference's pretraining loop (Model_Pretraining.py:80-106) on a fixed batch"""
    import Models as M
    torch.manual_seed(0); random.seed(0)
    model = M.HSIMAE(**TINY).to(DEV)
    no_decay = ["bias", "norm"]
    groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
              {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=5e-3, betas=(0.9, 0.95))
    x = torch.randn(256, 1, 32, 9, 9, device=DEV)
    losses = []
    for _ in range(60):
        loss, _, _ = model(x, mask_ratio=0.5)
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.005, losses


def test_scene_classification_matches_cube_batches_and_oracle():
    """dense per-pixel inference (BASELINE config 5): on-device sliding windows == host-materialised cubes"""
    import Models as M
    from hsimae_b200.scene import classify_scene, symmetric_pad_hwc
    g = tiny_geometry(17)
    sd = O.make_state(g, seed=21, decoder=False, head=True, randomize_affine=True)
    kw = {k: v for k, v in dict(TINY, num_class=17).items() if not k.startswith("decoder") and k != "norm_pix_loss"}
    vit = M.HSIViT(**kw)
    vit.load_state_dict({**vit.state_dict(), **sd})
    vit = vit.to(DEV).eval()
    torch.manual_seed(2)
    H, W = 13, 11
    scene = torch.randn(H, W, 32)
    logits = classify_scene(vit, scene.to(DEV), batch=50)          # several partial batches
    assert logits.shape == (H * W, 17)
    # the reference's host path: symmetric padding + one cube per pixel (Utils/Preprocessing.py:205-213), HWC -> [1,C,H,W]
    padded = np.pad(scene.numpy(), ((4, 4), (4, 4), (0, 0)), "symmetric")
    assert np.array_equal(symmetric_pad_hwc(scene, 4).numpy(), padded)
    cubes = np.stack([padded[r:r + 9, c:c + 9, :] for r in range(H) for c in range(W)])
    x = torch.from_numpy(cubes).permute(0, 3, 1, 2).unsqueeze(1).contiguous()
    with torch.no_grad():
        via_cubes = vit(x.to(DEV))
    assert torch.equal(logits, via_cubes)                            # same kernels, same values: bit-identical
    ref = O.vit_forward(sd, x, g)
    assert rel_err(logits, ref) < TOL_ACT
    assert (logits[:, 1:].argmax(1).cpu() == ref[:, 1:].argmax(1)).float().mean() > 0.97


def test_other_patch_geometry():
    """a geometry other than the reference's 9x9x32 / 3x3x8 exercises the generic (non-specialised) kernels"""
    import Models as M
    kw = dict(img_size=8, patch_size=2, in_chans=1, bands=12, b_patch_size=4, embed_dim=64, depth=3, num_heads=4, s_depth=2,
              decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=4, norm_pix_loss=True, trunc_init=True)
    g = O.Geometry(img_size=8, patch_size=2, bands=12, b_patch_size=4, embed_dim=64, depth=3, s_depth=2, num_heads=4,
                   decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=4)
    sd = O.make_state(g, seed=4, randomize_affine=True)
    model = M.HSIMAE(**kw)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV)
    torch.manual_seed(8); random.seed(8)
    x = torch.randn(9, 1, 12, 8, 8, device=DEV)
    loss, pred, mask = model(x, mask_ratio=0.6)
    loss.backward()
    aux = model._last
    out, grads = O.pretrain_step_grads(sd, x.cpu(), g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"])
    assert torch.equal(aux["ids_keep"].cpu(), out["ids_keep"]) and torch.equal(mask.cpu(), out["mask_img"])
    assert abs(loss.item() - out["loss"].item()) <= TOL_LOSS * abs(out["loss"].item())
    assert rel_err(pred, out["pred_img"]) < TOL_ACT
    _check_grads(model, grads)


def test_loss_curve_tracks_oracle_training():
    """The reference's pretraining loop (Model_Pretraining.py:80-106: AdamW lr 5e-3, betas (0.9, 0.95), decay split on
    'bias'/'norm') run side by side on the CUDA path and on the fp32 CPU oracle, fed the SAME noise every step: the two
    loss curves must stay within 2e-3 relative of each other over 12 optimiser steps (measured on B200: 2.5e-5; the
    bound is the single-step loss tolerance, Adam's sign-like early steps amplify rounding noise on near-zero gradients)."""
    import Models as M
    g = tiny_geometry()
    sd = O.make_state(g, seed=3)
    model = M.HSIMAE(**TINY)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in sd.items()}
    no_decay = ("bias", "norm")

    def groups(named):
        named = [(n, p) for n, p in named if p.requires_grad]
        return [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
                {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt_g = torch.optim.AdamW(groups(model.named_parameters()), lr=5e-3, betas=(0.9, 0.95))
    opt_c = torch.optim.AdamW(groups(leaves.items()), lr=5e-3, betas=(0.9, 0.95))
    torch.manual_seed(11); random.seed(11)
    x = torch.randn(96, 1, 32, 9, 9)
    xg = x.to(DEV)
    curve_g, curve_c = [], []
    for _ in range(12):
        loss, _, _ = model(xg, mask_ratio=0.5)
        opt_g.zero_grad(); loss.backward(); opt_g.step()
        aux = model._last
        out = O.pretrain_forward(leaves, x, g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"])
        assert torch.equal(aux["ids_keep"].cpu(), out["ids_keep"])
        opt_c.zero_grad(); out["loss"].backward(); opt_c.step()
        curve_g.append(loss.item()); curve_c.append(out["loss"].item())
    dev = max(abs(a - b) / abs(b) for a, b in zip(curve_g, curve_c))
    print("loss curves (cuda / oracle):", [f"{a:.4f}/{b:.4f}" for a, b in zip(curve_g, curve_c)], "max rel dev", dev)
    assert curve_c[-1] < curve_c[0] and curve_g[-1] < curve_g[0]
    assert dev < 2e-3, (curve_g, curve_c)


def test_packed_weights_follow_parameter_updates():
    """The bf16 operand arenas are a cache of the fp32 masters keyed by (storage pointer, version): every way the
    reference drivers change weights -- load_state_dict (Model_Finetuning.py:89-96), optimiser steps, .to() -- and
    replacing a Parameter object must be seen by the next forward."""
    import Models as M
    kw = {k: v for k, v in dict(TINY, num_class=5).items() if not k.startswith("decoder") and k != "norm_pix_loss"}
    g = tiny_geometry(5)
    torch.manual_seed(5)
    x = torch.randn(12, 1, 32, 9, 9, device=DEV)

    def fresh(sd):
        m = M.HSIViT(**kw)
        m.load_state_dict({**m.state_dict(), **sd})
        with torch.no_grad():
            return m.to(DEV).eval()(x)
    sd_a = O.make_state(g, seed=1, decoder=False, head=True)
    sd_b = O.make_state(g, seed=2, decoder=False, head=True)
    vit = M.HSIViT(**kw)
    vit.load_state_dict({**vit.state_dict(), **sd_a})
    vit = vit.to(DEV).eval()
    with torch.no_grad():
        ya = vit(x)
        assert torch.equal(ya, vit(x))                                   # cached arenas: same answer
        vit.load_state_dict({**vit.state_dict(), **{k: v.to(DEV) for k, v in sd_b.items()}})
        yb = vit(x)
    assert torch.equal(yb, fresh(sd_b)) and not torch.equal(ya, yb)
    # in-place update under no_grad (what optimisers do)
    with torch.no_grad():
        vit.cls_head.weight.mul_(2.0); vit.cls_head.bias.mul_(2.0)
        assert torch.allclose(vit(x), 2.0 * yb, rtol=1e-5, atol=1e-6)
        # a replaced Parameter object
        vit.cls_head.weight = torch.nn.Parameter(vit.cls_head.weight.detach() * 0.5)
        vit.cls_head.bias = torch.nn.Parameter(vit.cls_head.bias.detach() * 0.5)
        assert torch.allclose(vit(x), yb, rtol=1e-5, atol=1e-6)
        # dtype round trip re-allocates every storage
        vit = vit.double().float()
        assert torch.allclose(vit(x), yb, rtol=1e-5, atol=1e-6)
        # a write through the `.data` alias is invisible to torch's version counter: explicit invalidation
        vit.cls_head.weight.data.mul_(2.0); vit.cls_head.bias.data.mul_(2.0)
        vit.invalidate_weight_cache()
        assert torch.allclose(vit(x), 2.0 * yb, rtol=1e-5, atol=1e-6)
        vit.cls_head.weight.mul_(0.5); vit.cls_head.bias.mul_(0.5)
    # wrong dtype / device are rejected loudly, not silently reinterpreted
    vit.cls_head.weight = torch.nn.Parameter(vit.cls_head.weight.detach().double())
    with pytest.raises(RuntimeError):
        vit(x)
    # a deep copy owns its own runtime and parameters
    import copy
    vit.cls_head.weight = torch.nn.Parameter(vit.cls_head.weight.detach().float())
    twin = copy.deepcopy(vit)
    with torch.no_grad():
        twin.cls_head.weight.mul_(3.0); twin.cls_head.bias.mul_(3.0)
        assert torch.allclose(twin(x), 3.0 * yb, rtol=1e-5, atol=1e-6)
        assert torch.allclose(vit(x), yb, rtol=1e-5, atol=1e-6)


def test_headline_config_matches_oracle_at_full_size():
    """BASELINE config-2 itself -- HSIMAE-Large, N = 4096, mask 0.5 -- against the CPU fp32 oracle on the same weights,
    batch and noise (one oracle step at this size is seconds of host time): loss, the pixel reconstruction and six
    named gradient tensors spanning the path (decoder head, decoder, fusion, spectral, spatial, patch embedding)."""
    import Models as M
    kw = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=256, depth=12, num_heads=16, s_depth=9,
              decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
    g = O.Geometry(embed_dim=256, num_heads=16)
    sd = O.make_state(g, seed=4096, randomize_affine=True)
    model = M.HSIMAE(**kw)
    model.load_state_dict({**model.state_dict(), **sd})
    model = model.to(DEV)
    torch.manual_seed(7); random.seed(7)
    x = torch.randn(4096, 1, 32, 9, 9, device=DEV)
    loss, pred, mask = model(x, mask_ratio=0.5)
    loss.backward()
    aux = model._last
    torch.set_num_threads(max(1, (__import__("os").cpu_count() or 1)))
    out, grads = O.pretrain_step_grads(sd, x.cpu(), g, aux["noise_t"].cpu(), aux["noise_l"].cpu(), aux["lt"], aux["ll"])
    assert torch.equal(aux["ids_keep"].cpu(), out["ids_keep"]) and torch.equal(mask.cpu(), out["mask_img"])
    assert abs(loss.item() - out["loss"].item()) <= TOL_LOSS * abs(out["loss"].item())
    assert rel_err(pred, out["pred_img"]) < TOL_ACT
    named = dict(model.named_parameters())
    for k in ("decoder_pred.weight", "decoder_blocks.3.mlp.w2.weight", "blocks.1.attn.proj.weight", "blocks_2.4.mlp.w1.weight",
              "blocks_1.0.attn.q.weight", "patch_embed.proj.weight"):
        e = rel_err(named[k].grad, grads[k])
        assert e < TOL_GRAD, f"{k}: {e:.3g}"
