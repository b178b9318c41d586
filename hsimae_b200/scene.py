"""Dense per-pixel classification of a whole hyperspectral scene (BASELINE.json configs[4]).

The reference materialises one 9x9xC cube per pixel on the host (`Utils/Preprocessing.py:205-213`:
symmetric padding + `splitHSI` with unit step, ~1.15 GB for Salinas) and feeds them to `HSIViT` in batches of 256
(`Model_Finetuning.py:264-278`).  Here the padded scene (a few MB) stays in HBM and the patch-embedding kernel
gathers each window on the fly, so the scene is classified in a handful of large launches.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .modules import _ptr, _stream


def symmetric_pad_hwc(scene: torch.Tensor, pad: int) -> torch.Tensor:
    """numpy.pad(..., 'symmetric') over the two spatial axes of an [H, W, C] tensor (edge sample repeated)."""
    H, W = scene.shape[0], scene.shape[1]
    if pad > H or pad > W:
        raise ValueError("scene smaller than the padding")

    def index(n):
        a = torch.arange(n, device=scene.device)
        return torch.cat([a[:pad].flip(0), a, a[n - pad:].flip(0)])

    return scene.index_select(0, index(H)).index_select(1, index(W)).contiguous()


@torch.no_grad()
def classify_scene(model, scene: torch.Tensor, batch: int = 8192, pad: bool = True) -> torch.Tensor:
    """logits [H*W, num_class] for every pixel of `scene` ([H, W, bands], fp32), pixel-centred windows.

    `model` is an `HSIViT` (or a `DualViT`, whose encoder + head are used) in eval mode on a CUDA device."""
    if scene.dim() != 3:
        raise ValueError("scene must be [H, W, bands]")
    rt, _ = model._prepare(scene if scene.is_cuda else scene.to(next(model.parameters()).device))
    dev = rt.device
    pe = model.patch_embed
    if scene.shape[2] != pe.bands:
        raise ValueError(f"scene has {scene.shape[2]} bands, model expects {pe.bands}")
    if model.training:
        raise RuntimeError("classify_scene is an inference path: call model.eval() first")
    img = pe.img_size[0]
    x = scene.to(dev, torch.float32)
    x = symmetric_pad_hwc(x, img // 2) if pad else x.contiguous()
    Hp, Wp = x.shape[0], x.shape[1]
    Ho, Wo = Hp - img + 1, Wp - img + 1
    total = Ho * Wo
    T, Lp = pe.b_grid_size, pe.grid_size ** 2
    ncls = model.cls_head.out_features
    out = torch.empty(total, ncls, dtype=torch.float32, device=dev)
    ws = rt.enc_ws(min(batch, total), T, Lp, False, dev)
    pooled = torch.empty(min(batch, total), T * model.dim, dtype=torch.float32, device=dev)
    for p0 in range(0, total, batch):
        n = min(batch, total - p0)
        _lib.check(rt.lib.hsimae_encoder_forward_scene(rt.plan, _ptr(rt.wb), _ptr(rt.wf), _ptr(x), Hp, Wp, p0, n, _ptr(ws), ws.numel(),
                                                       _stream()), "encoder_forward_scene")
        _lib.check(rt.lib.hsimae_head_forward(rt.plan, _ptr(rt.wf), n, _ptr(ws), 0, _ptr(pooled), _ptr(out[p0:p0 + n]), _stream()),
                   "head_forward")
    return out
