"""CPU fp32 oracle for the HSIMAE hot path -- TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *functional* restatement (pure functions over a
``state_dict``-style mapping of fp32 tensors) of the algorithm implemented by
the reference's ``Models.py``.  It exists so that the CUDA path can be checked
on machines where ``/root/reference`` is absent (the GPU box).  Nothing in the
product path (``Models.py``, ``hsimae_b200/``) may import it: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do.

Parity pin: the reference ships no tests or golden vectors,
so this restatement is pinned against the *live* reference imported from
``/root/reference`` -- see ``oracle/make_golden.py`` (generator, run in the build
container) and ``tests/test_oracle_vs_reference.py`` (skipped when the
reference is absent) -- and against the committed fixtures in ``tests/golden/``
that the generator wrote from the reference's outputs.

Every function cites the reference lines (``Models.py:<line>``) it follows.
Backward passes are obtained from torch autograd over these functions, exactly
as the reference obtains its own (``Model_Pretraining.py:101``).
"""
from __future__ import annotations

import math
import random as _pyrandom
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass
class Geometry:
    """Static shape facts of one model instance (Models.py:107-149, 312-352)."""

    img_size: int = 9
    patch_size: int = 3
    bands: int = 32
    b_patch_size: int = 8
    embed_dim: int = 256
    depth: int = 12
    s_depth: int = 9
    num_heads: int = 16
    decoder_embed_dim: int = 64
    decoder_depth: int = 8
    decoder_num_heads: int = 8
    mlp_ratio: float = 4.0
    num_class: int = 0

    @property
    def T(self) -> int:  # spectral groups
        return self.bands // self.b_patch_size

    @property
    def G(self) -> int:  # spatial grid side
        return self.img_size // self.patch_size

    @property
    def L(self) -> int:  # spatial positions
        return self.G * self.G

    @property
    def P(self) -> int:  # tokens per sample
        return self.T * self.L

    @property
    def patch_dim(self) -> int:
        return self.b_patch_size * self.patch_size * self.patch_size

    @property
    def n_fusion(self) -> int:
        # Models.py:385-398 -- literal 12, count = depth - s_depth
        return max(self.depth - self.s_depth, 0) if self.s_depth < 12 else 0


def swiglu_hidden(dim: int, mlp_ratio: float = 4.0) -> int:
    """Hidden width of the gated MLP (Models.py:225 with the arguments Block
    passes at Models.py:300-301: hidden=int(dim*ratio), multiple_of=ratio)."""
    hidden = int(dim * mlp_ratio)
    m = mlp_ratio
    return int(m * ((2 * hidden // 3 + m - 1) // m))


# --------------------------------------------------------------------------
# fixed sin-cos position table (Models.py:11-101)
# --------------------------------------------------------------------------
def _sincos_1d(width: int, pos: np.ndarray) -> np.ndarray:
    """Models.py:86-101: [sin | cos] of pos * 10000^(-i/(width/2))."""
    half = width // 2
    freq = np.arange(half, dtype=np.float32)
    freq /= width / 2.0
    freq = 1.0 / 10000 ** freq
    ang = np.einsum("m,d->md", pos.reshape(-1), freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_table(width: int, T: int, G: int) -> Tensor:
    """[1, T*G*G, width] table: first half encodes the spectral index, second
    half the (h, w) grid position (Models.py:11-47, 71-83)."""
    assert width % 4 == 0
    half = width // 2
    gw, gh = np.meshgrid(np.arange(G, dtype=np.float32), np.arange(G, dtype=np.float32))
    # reference meshgrid(grid_w, grid_h) -> grid[0] varies along w  (Models.py:19-22)
    spatial = np.concatenate([_sincos_1d(half // 2, gw), _sincos_1d(half // 2, gh)], axis=1)
    spectral = _sincos_1d(half, np.arange(T, dtype=np.float32))
    spectral = np.repeat(spectral[:, None, :], G * G, axis=1)
    spatial = np.repeat(spatial[None, :, :], T, axis=0)
    table = np.concatenate([spectral, spatial], axis=-1).reshape(-1, width)
    return torch.tensor(table, dtype=torch.float32).unsqueeze(0)


# --------------------------------------------------------------------------
# masking (Models.py:484-535)
# --------------------------------------------------------------------------
def choose_visible_shape(T: int, L: int, mask_ratio: float, rng=_pyrandom) -> Tuple[int, int]:
    """(len_t, len_l) in [2..T]x[2..L] whose product is closest to the kept
    token count; ties broken by one ``random.sample`` draw that is consumed even
    when there is a single candidate (Models.py:484-493)."""
    cands = [(t, l) for t in range(2, T + 1) for l in range(2, L + 1)]
    keep = (1 - mask_ratio) * T * L
    # the reference compares in float32 tensor arithmetic (Models.py:486-489)
    prod = torch.tensor([t * l for t, l in cands])
    diff = abs(keep - prod)
    best = torch.where(diff == torch.min(diff))[0]
    pick = rng.sample(range(len(best)), 1)[0]
    t, l = cands[int(best[pick])]
    return int(t), int(l)


def _stable_rank(noise: Tensor) -> Tensor:
    """rank[n, i] = #{j : noise[n,j] < noise[n,i] or (== and j < i)}."""
    a = noise.unsqueeze(2)  # [N, n, 1]  (i)
    b = noise.unsqueeze(1)  # [N, 1, n]  (j)
    n = noise.shape[1]
    idx = torch.arange(n, device=noise.device)
    earlier = (idx.view(1, 1, n) < idx.view(1, n, 1))
    return ((b < a) | ((b == a) & earlier)).sum(dim=2)


def structured_mask(noise_t: Tensor, noise_l: Tensor, len_t: int, len_l: int):
    """Rank-based restatement of Models.py:495-535.

    The reference keeps the ``len_t`` spectral groups with the smallest
    ``noise_t`` and the ``len_l`` spatial positions with the smallest
    ``noise_l``; a token is visible iff both its group and its position are
    kept.  Its argsort-of-(mask_1+mask_2+linspace(0,0.5)) (Models.py:520-524)
    orders tokens by (number of dropped axes, raster index), so:

      ids_shuffle = [visible tokens in raster order,
                     tokens dropped on exactly one axis in raster order,
                     tokens dropped on both axes in raster order]

    Exact float ties between *distinct* noise draws are undefined in the
    reference (unstable argsort); they are resolved lowest-index-first here and
    in the CUDA kernel.
    Returns ids_keep [N,len_t*len_l] i64, ids_restore [N,T*L] i64, mask [N,T*L] f32.
    """
    N, T = noise_t.shape
    L = noise_l.shape[1]
    keep_t = _stable_rank(noise_t) < len_t  # [N,T]
    keep_l = _stable_rank(noise_l) < len_l  # [N,L]
    dropped = (~keep_t).long().unsqueeze(2) + (~keep_l).long().unsqueeze(1)  # [N,T,L] in {0,1,2}
    dropped = dropped.reshape(N, T * L)
    key = dropped * (T * L) + torch.arange(T * L, device=noise_t.device).unsqueeze(0)
    ids_shuffle = torch.argsort(key, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, : len_t * len_l]
    mask = (dropped > 0).to(torch.float32)
    return ids_keep, ids_restore, mask


# --------------------------------------------------------------------------
# patch <-> cube index permutations (Models.py:461-482)
# --------------------------------------------------------------------------
def cube_to_patches(imgs: Tensor, g: Geometry) -> Tensor:
    """[N,1,bands,H,W] -> [N, T*L, u*p*p]; inner order (u, p, q), token order
    (t, h, w)  (Models.py:461-473)."""
    N = imgs.shape[0]
    u, p, T, G = g.b_patch_size, g.patch_size, g.T, g.G
    x = imgs.reshape(N, T, u, G, p, G, p)
    x = x.permute(0, 1, 3, 5, 2, 4, 6)  # n t h w u p q
    return x.reshape(N, T * G * G, u * p * p)


def patches_to_cube(x: Tensor, g: Geometry) -> Tensor:
    """Inverse of :func:`cube_to_patches` (Models.py:475-482)."""
    N = x.shape[0]
    u, p, T, G = g.b_patch_size, g.patch_size, g.T, g.G
    x = x.reshape(N, T, G, G, u, p, p)
    x = x.permute(0, 1, 4, 2, 5, 3, 6)  # n t u h p w q
    return x.reshape(N, 1, T * u, G * p, G * p)


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def patch_embed(sd: Dict[str, Tensor], imgs: Tensor, g: Geometry) -> Tensor:
    """Conv3d with stride == kernel is a per-patch linear map
    (Models.py:147-149, 157-158) -> [N, T*L, D]."""
    w = sd["patch_embed.proj.weight"]
    b = sd["patch_embed.proj.bias"]
    return F.linear(cube_to_patches(imgs, g), w.reshape(w.shape[0], -1), b)


def attention(sd, pre: str, x: Tensor, heads: int) -> Tensor:
    """Models.py:192-219: separate q/k/v projections, softmax(q k^T / sqrt(hd)) v,
    output projection."""
    B, S, C = x.shape
    hd = C // heads

    def split(name):
        y = F.linear(x, sd[pre + name + ".weight"], sd.get(pre + name + ".bias"))
        return y.reshape(B, S, heads, hd).transpose(1, 2)

    q, k, v = split("q"), split("k"), split("v")
    att = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    y = (att @ v).transpose(1, 2).reshape(B, S, C)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def gated_mlp(sd, pre: str, x: Tensor) -> Tensor:
    """Models.py:231-232: w2(silu(w1 x) * w3 x)."""
    a = F.linear(x, sd[pre + "w1.weight"], sd[pre + "w1.bias"])
    b = F.linear(x, sd[pre + "w3.weight"], sd[pre + "w3.bias"])
    return F.linear(F.silu(a) * b, sd[pre + "w2.weight"], sd[pre + "w2.bias"])


def block(sd, pre: str, x: Tensor, heads: int, keep1: Optional[Tensor] = None,
          keep2: Optional[Tensor] = None) -> Tensor:
    """Pre-LN transformer block (Models.py:303-306).  ``keep1``/``keep2`` are the
    already-scaled stochastic-depth factors of shape [rows,1,1]
    (Models.py:246-251); ``None`` means Identity."""
    C = x.shape[-1]
    h = attention(sd, pre + "attn.", F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5), heads)
    x = x + (h if keep1 is None else h * keep1)
    h = gated_mlp(sd, pre + "mlp.", F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5))
    x = x + (h if keep2 is None else h * keep2)
    return x


def _split_encoders(sd, x: Tensor, g: Geometry, lt: int, ll: int, drops=None) -> Tensor:
    """Spatial encoder over '(b t) l c', spectral encoder over '(b l) t c', summed,
    then fusion blocks and the final norm (Models.py:552-570).  ``drops`` is an
    optional dict {(stack, i, 1|2): factor tensor}."""
    N, _, C = x.shape
    drops = drops or {}
    if g.s_depth > 0:
        x4 = x.reshape(N, lt, ll, C)
        xs = x4.reshape(N * lt, ll, C)
        xc = x4.transpose(1, 2).reshape(N * ll, lt, C)
        for i in range(g.s_depth):
            xs = block(sd, f"blocks_1.{i}.", xs, g.num_heads, drops.get((1, i, 1)), drops.get((1, i, 2)))
        for i in range(g.s_depth):
            xc = block(sd, f"blocks_2.{i}.", xc, g.num_heads, drops.get((2, i, 1)), drops.get((2, i, 2)))
        x = xs.reshape(N, lt * ll, C) + xc.reshape(N, ll, lt, C).transpose(1, 2).reshape(N, lt * ll, C)
    for i in range(g.n_fusion):
        x = block(sd, f"blocks.{i}.", x, g.num_heads, drops.get((0, i, 1)), drops.get((0, i, 2)))
    return F.layer_norm(x, (C,), sd["norm.weight"], sd["norm.bias"], 1e-5)


def encode_masked(sd, imgs: Tensor, g: Geometry, ids_keep: Tensor, lt: int, ll: int, drops=None) -> Tensor:
    """Models.py:537-571 (HSIMAE.forward_encoder) / :896-923 (DualViT.forward_mask_encoder)
    given the mask indices."""
    x = patch_embed(sd, imgs, g)
    D = x.shape[-1]
    idx = ids_keep.unsqueeze(-1).expand(-1, -1, D)
    x = torch.gather(x, 1, idx) + torch.gather(sd["pos_embed"].expand(x.shape[0], -1, -1), 1, idx)
    return _split_encoders(sd, x, g, lt, ll, drops)


def encode_full(sd, imgs: Tensor, g: Geometry, drops=None) -> Tensor:
    """Unmasked encoder (Models.py:869-894, 1119-1145)."""
    x = patch_embed(sd, imgs, g) + sd["pos_embed"]
    return _split_encoders(sd, x, g, g.T, g.L, drops)


def classify(sd, latent: Tensor, g: Geometry) -> Tensor:
    """'AGG' head: [N,T,L,C] -> [N,L,T*C], mean over L, linear
    (Models.py:964-973, 1147-1156)."""
    N, _, C = latent.shape
    z = latent.reshape(N, g.T, g.L, C).permute(0, 2, 1, 3).reshape(N, g.L, g.T * C).mean(1)
    return F.linear(z, sd["cls_head.weight"], sd["cls_head.bias"])


def decode(sd, latent: Tensor, ids_restore: Tensor, g: Geometry) -> Tensor:
    """Models.py:573-601.  The masked slots are filled with the per-sample mean
    of the visible decoder-embedded tokens (Models.py:583-584); the learned
    ``mask_token`` parameter is never read."""
    y = F.linear(latent, sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    N, K, C = y.shape
    fill = y.mean(1, keepdim=True).expand(N, g.P - K, C)
    y = torch.gather(torch.cat([y, fill], 1), 1, ids_restore.unsqueeze(-1).expand(-1, -1, C))
    y = y + sd["decoder_pos_embed"]
    for i in range(g.decoder_depth):
        y = block(sd, f"decoder_blocks.{i}.", y, g.decoder_num_heads)
    y = F.layer_norm(y, (C,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], 1e-5)
    return F.linear(y, sd["decoder_pred.weight"], sd["decoder_pred.bias"])


def reconstruction_loss(imgs: Tensor, pred: Tensor, mask: Tensor, g: Geometry, norm_pix: bool = True):
    """Models.py:603-616.  Unbiased per-patch variance, eps added to the variance.
    Returns (loss, mean, std) with std = sqrt(var+1e-6) as stashed at :609-610."""
    tgt = cube_to_patches(imgs, g)
    mean = std = None
    if norm_pix:
        mean = tgt.mean(-1, keepdim=True)
        std = (tgt.var(-1, keepdim=True) + 1.0e-6) ** 0.5
        tgt = (tgt - mean) / std
    per_patch = ((pred - tgt) ** 2).mean(-1)
    return (per_patch * mask).sum() / mask.sum(), mean, std


def to_pixels(pred: Tensor, mask: Tensor, mean, std, g: Geometry):
    """Models.py:618-625: de-normalise and un-patchify pred; broadcast mask."""
    m = patches_to_cube(mask.unsqueeze(2).repeat(1, 1, pred.shape[2]), g)
    if mean is not None:
        pred = pred * std + mean
    return patches_to_cube(pred, g), m


# --------------------------------------------------------------------------
# whole-model entry points
# --------------------------------------------------------------------------
def pretrain_forward(sd, imgs: Tensor, g: Geometry, noise_t: Tensor, noise_l: Tensor,
                     lt: int, ll: int, norm_pix: bool = True):
    """HSIMAE.forward (Models.py:627-634) with the random draws supplied by the
    caller.  Returns dict(loss, pred_img, mask_img, pred, latent, ids_keep,
    ids_restore, mask)."""
    ids_keep, ids_restore, mask = structured_mask(noise_t, noise_l, lt, ll)
    latent = encode_masked(sd, imgs, g, ids_keep, lt, ll)
    pred = decode(sd, latent, ids_restore, g)
    loss, mean, std = reconstruction_loss(imgs, pred, mask, g, norm_pix)
    pred_img, mask_img = to_pixels(pred, mask, mean, std, g)
    return dict(loss=loss, pred_img=pred_img, mask_img=mask_img, pred=pred, latent=latent,
                ids_keep=ids_keep, ids_restore=ids_restore, mask=mask)


def dual_forward(sd, imgs: Tensor, imgs_u: Optional[Tensor], g: Geometry, noise_t=None, noise_l=None,
                 lt: int = 0, ll: int = 0, drops_full=None, drops_masked=None, norm_pix: bool = True):
    """DualViT.forward (Models.py:975-993)."""
    logits = classify(sd, encode_full(sd, imgs, g, drops_full), g)
    if imgs_u is None:
        return dict(logits=logits)
    both = torch.cat([imgs, imgs_u], 0)
    ids_keep, ids_restore, mask = structured_mask(noise_t, noise_l, lt, ll)
    latent = encode_masked(sd, both, g, ids_keep, lt, ll, drops_masked)
    pred = decode(sd, latent, ids_restore, g)
    loss, mean, std = reconstruction_loss(both, pred, mask, g, norm_pix)
    pred_img, mask_img = to_pixels(pred, mask, mean, std, g)
    return dict(logits=logits, loss=loss, pred_img=pred_img, mask_img=mask_img, pred=pred,
                ids_keep=ids_keep, ids_restore=ids_restore, mask=mask)


def vit_forward(sd, imgs: Tensor, g: Geometry, drops=None) -> Tensor:
    """HSIViT.forward (Models.py:1158-1160)."""
    return classify(sd, encode_full(sd, imgs, g, drops), g)


# --------------------------------------------------------------------------
# parameter factory (names/shapes of Models.py:342-424, 736; init :429-459)
# --------------------------------------------------------------------------
def _block_params(sd, pre: str, dim: int, ratio: float, gen):
    hid = swiglu_hidden(dim, ratio)

    def lin(name, out_f, in_f):
        sd[pre + name + ".weight"] = torch.nn.init.trunc_normal_(torch.empty(out_f, in_f), std=0.02, generator=gen)
        sd[pre + name + ".bias"] = torch.zeros(out_f)

    sd[pre + "norm1.weight"], sd[pre + "norm1.bias"] = torch.ones(dim), torch.zeros(dim)
    for n in ("q", "k", "v", "proj"):
        lin("attn." + n, dim, dim)
    sd[pre + "norm2.weight"], sd[pre + "norm2.bias"] = torch.ones(dim), torch.zeros(dim)
    lin("mlp.w1", hid, dim)
    lin("mlp.w2", dim, hid)
    lin("mlp.w3", hid, dim)


def make_state(g: Geometry, seed: int = 0, decoder: bool = True, head: bool = False,
               randomize_affine: bool = False) -> Dict[str, Tensor]:
    """Random fp32 parameters under the reference's state_dict names, with the
    reference's init statistics (trunc-normal std 1 for the conv, 0.02 for
    Linears; Models.py:437-459).  ``randomize_affine`` additionally perturbs
    biases / LayerNorm affine so that tests exercise them."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    D, Dd = g.embed_dim, g.decoder_embed_dim
    sd["pos_embed"] = sincos_table(D, g.T, g.G)
    sd["patch_embed.proj.weight"] = torch.nn.init.trunc_normal_(
        torch.empty(D, 1, g.b_patch_size, g.patch_size, g.patch_size), generator=gen)
    sd["patch_embed.proj.bias"] = torch.zeros(D)
    for i in range(g.s_depth):
        _block_params(sd, f"blocks_1.{i}.", D, g.mlp_ratio, gen)
    for i in range(g.s_depth):
        _block_params(sd, f"blocks_2.{i}.", D, g.mlp_ratio, gen)
    for i in range(g.n_fusion):
        _block_params(sd, f"blocks.{i}.", D, g.mlp_ratio, gen)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(D), torch.zeros(D)
    if head:
        sd["cls_head.weight"] = torch.nn.init.trunc_normal_(torch.empty(g.num_class, D * g.T), std=0.02, generator=gen)
        sd["cls_head.bias"] = torch.zeros(g.num_class)
    if decoder:
        sd["mask_token"] = torch.zeros(1, 1, Dd)
        sd["decoder_embed.weight"] = torch.nn.init.trunc_normal_(torch.empty(Dd, D), std=0.02, generator=gen)
        sd["decoder_embed.bias"] = torch.zeros(Dd)
        sd["decoder_pos_embed"] = sincos_table(Dd, g.T, g.G)
        for i in range(g.decoder_depth):
            _block_params(sd, f"decoder_blocks.{i}.", Dd, g.mlp_ratio, gen)
        sd["decoder_norm.weight"], sd["decoder_norm.bias"] = torch.ones(Dd), torch.zeros(Dd)
        sd["decoder_pred.weight"] = torch.nn.init.trunc_normal_(torch.empty(g.patch_dim, Dd), std=0.02, generator=gen)
        sd["decoder_pred.bias"] = torch.zeros(g.patch_dim)
    if randomize_affine:
        for k, v in sd.items():
            if k in ("pos_embed", "decoder_pos_embed", "mask_token"):
                continue
            if k.endswith(".bias"):
                v.copy_(0.05 * torch.randn(v.shape, generator=gen))
            elif "norm" in k and k.endswith(".weight"):
                v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=gen))
    return sd


FROZEN = ("pos_embed", "decoder_pos_embed", "mask_token")


def pretrain_step_grads(sd, imgs, g: Geometry, noise_t, noise_l, lt, ll):
    """loss + {name: grad} for one HSIMAE fwd+bwd; the frozen/unused tensors
    (Models.py:434-435 and the never-read mask_token) get no gradient."""
    leaves = {k: v.detach().clone().requires_grad_(k not in FROZEN) for k, v in sd.items()}
    out = pretrain_forward(leaves, imgs, g, noise_t, noise_l, lt, ll)
    out["loss"].backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return out, grads
