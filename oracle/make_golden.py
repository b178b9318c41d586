"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference/Models.py).

Run in the build container (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin (a) the oracle restatement and (b) the CUDA path against what
the reference itself computes on CPU in fp32.  TEST INFRASTRUCTURE ONLY.

Files written
  tiny_pretrain.npz : HSIMAE (dim 64, heads 4, depth 3, s_depth 2, decoder 32x1x4), N=8, mask 0.5 --
                      full state_dict, input, noise, visible shape, loss, pred, mask, ids, latent, all gradients.
  tiny_dual.npz     : DualViT with the same encoder (+ 17-way head), drop_path=0: train-mode tuple + gradients,
                      eval-mode logits; HSIViT logits from the shared keys.
  kat_masks.npz     : Base/Large known-answer masks of BASELINE.md section 5 (noise from torch.manual_seed(7)).
  init_hashes.npz   : SHA-1 of every state_dict tensor after seeded construction (checks init/RNG-order parity).
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib.util
import io
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("_reference_models", "/root/reference/Models.py")
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


TINY = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=64, depth=3, num_heads=4, s_depth=2,
            decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=4, norm_pix_loss=True, trunc_init=True)


def randomize(model, gen):
    """make biases / LayerNorm affine non-trivial so the fixtures exercise them"""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n in ("pos_embed", "decoder_pos_embed", "mask_token"):
                continue
            if n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))
            elif "norm" in n and n.endswith(".weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen))


def np_state(model):
    return {"sd/" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    torch.set_num_threads(4)

    # ---------------- tiny pretrain ----------------
    torch.manual_seed(1234); random.seed(1234)
    m = quiet(R.HSIMAE, **TINY)
    randomize(m, torch.Generator().manual_seed(5))
    torch.manual_seed(99); random.seed(99)
    x = torch.randn(8, 1, 32, 9, 9)
    rng = (torch.get_rng_state(), random.getstate())
    loss, pred, mask = m(x, mask_ratio=0.5)
    loss.backward()
    lt, ll = int(m.len_t), int(m.len_l)
    torch.set_rng_state(rng[0]); random.setstate(rng[1])
    random.sample(range(2), 1)  # the visible-shape draw (two candidates at ratio 0.5)
    noise_t = torch.rand(8, 4); noise_l = torch.rand(8, 9)
    torch.set_rng_state(rng[0]); random.setstate(rng[1])
    latent, mask_tok, ids_restore, ids_keep = m.forward_encoder(x, 0.5)
    assert (int(m.len_t), int(m.len_l)) == (lt, ll)
    out = np_state(m)
    out.update(x=x.numpy(), noise_t=noise_t.numpy(), noise_l=noise_l.numpy(), len_t=lt, len_l=ll, loss=loss.item(),
               pred=pred.detach().numpy(), mask=mask.numpy(), latent=latent.detach().numpy(), mask_tokens=mask_tok.numpy(),
               ids_restore=ids_restore.numpy(), ids_keep=ids_keep.numpy())
    for n, p in m.named_parameters():
        if p.grad is not None:
            out["grad/" + n] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "tiny_pretrain.npz"), **out)
    print("tiny_pretrain: loss", loss.item(), "shape", (lt, ll))

    # ---------------- tiny dual / vit ----------------
    kw = dict(TINY); kw.update(num_class=17, drop_path=0.0)
    torch.manual_seed(4321); random.seed(4321)
    d = quiet(R.DualViT, **kw)
    randomize(d, torch.Generator().manual_seed(6))
    d.train()
    torch.manual_seed(77); random.seed(77)
    xl = torch.randn(6, 1, 32, 9, 9); xu = torch.randn(10, 1, 32, 9, 9)
    labels = torch.tensor([1, 5, 0, 16, 3, 9])
    rng = (torch.get_rng_state(), random.getstate())
    loss_rec, pred_rec, mask, logits = d(xl, xu, mask_ratio=0.8)
    total = 10.0 * loss_rec + torch.nn.functional.cross_entropy(logits, labels, ignore_index=0)
    total.backward()
    lt, ll = int(d.len_t), int(d.len_l)
    torch.set_rng_state(rng[0]); random.setstate(rng[1])
    random.sample(range(2), 1)
    noise_t = torch.rand(16, 4); noise_l = torch.rand(16, 9)
    d.eval()
    with torch.no_grad():
        logits_eval = d(xl)
    out = np_state(d)
    out.update(xl=xl.numpy(), xu=xu.numpy(), labels=labels.numpy(), noise_t=noise_t.numpy(), noise_l=noise_l.numpy(), len_t=lt,
               len_l=ll, loss_rec=loss_rec.item(), total=total.item(), pred_rec=pred_rec.detach().numpy(), mask=mask.numpy(),
               logits=logits.detach().numpy(), logits_eval=logits_eval.numpy())
    for n, p in d.named_parameters():
        if p.grad is not None:
            out["grad/" + n] = p.grad.numpy().copy()
    vkw = {k: v for k, v in kw.items() if not k.startswith("decoder") and k not in ("norm_pix_loss",)}
    v = quiet(R.HSIViT, **vkw)
    vd = v.state_dict()
    vd.update({k: t for k, t in d.state_dict().items() if k in vd})
    v.load_state_dict(vd); v.eval()
    with torch.no_grad():
        out["logits_vit"] = v(xl).numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_dual.npz"), **out)
    print("tiny_dual: loss_rec", loss_rec.item(), "total", total.item(), "shape", (lt, ll))

    # ---------------- Base/Large KAT masks ----------------
    kat = {}
    for name, dim in (("base", 128), ("large", 256)):
        torch.manual_seed(42); random.seed(42)
        big = quiet(R.HSIMAE, img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=dim, depth=12,
                    num_heads=dim // 16, s_depth=9, decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8,
                    norm_pix_loss=True, trunc_init=True)
        torch.manual_seed(7); random.seed(7)
        xb = torch.randn(64, 1, 32, 9, 9)
        rng = (torch.get_rng_state(), random.getstate())
        with torch.no_grad():
            latent, mk, idr, idk = big.forward_encoder(xb, 0.5)
        torch.set_rng_state(rng[0]); random.setstate(rng[1])
        with torch.no_grad():
            loss, _, _ = big(xb, mask_ratio=0.5)
        kat[name + "_ids_keep"] = idk.numpy(); kat[name + "_ids_restore"] = idr.numpy(); kat[name + "_mask"] = mk.numpy()
        kat[name + "_shape"] = np.array([int(big.len_t), int(big.len_l)])
        kat[name + "_loss"] = loss.item()
        kat[name + "_latent_absmean"] = latent.abs().mean().item()
        print(name, "loss", loss.item(), "ids_keep sha", hashlib.sha1(idk.numpy().tobytes()).hexdigest()[:16])
    torch.manual_seed(7)
    torch.randn(64, 1, 32, 9, 9)
    kat["noise_t"] = torch.rand(64, 4).numpy(); kat["noise_l"] = torch.rand(64, 9).numpy()
    np.savez_compressed(os.path.join(OUT, "kat_masks.npz"), **kat)

    # ---------------- init hashes ----------------
    hashes = {}
    for cls, extra in (("HSIMAE", {}), ("DualViT", dict(num_class=17, drop_path=0.2)), ("HSIViT", dict(num_class=17))):
        k = dict(TINY); k.update(extra)
        if cls == "HSIViT":
            k = {a: b for a, b in k.items() if not a.startswith("decoder") and a != "norm_pix_loss"}
        torch.manual_seed(42); random.seed(42)
        mod = quiet(getattr(R, cls), **k)
        h = hashlib.sha1()
        for key, t in mod.state_dict().items():
            h.update(key.encode()); h.update(t.numpy().tobytes())
        hashes[cls] = h.hexdigest()
        hashes[cls + "_keys"] = "\n".join(mod.state_dict().keys())
    np.savez_compressed(os.path.join(OUT, "init_hashes.npz"), **hashes)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
