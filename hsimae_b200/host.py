"""Host-side pieces of the hot path that are not device work: the visible-shape
choice, the fixed position tables, the stochastic-depth draws.  They follow the
reference's random-number consumption order exactly so that a seeded run of the
unchanged drivers produces the same masks.
"""
from __future__ import annotations

import random
from typing import List, Optional, Tuple

import numpy as np
import torch


def swiglu_hidden(dim: int, mlp_ratio: float = 4.0) -> int:
    """Hidden width of the gated MLP as the reference computes it
    (/root/reference/Models.py:225 called from :300-301)."""
    hidden = int(dim * mlp_ratio)
    m = mlp_ratio
    return int(m * ((2 * hidden // 3 + m - 1) // m))


def choose_visible_shape(T: int, L: int, mask_ratio: float) -> Tuple[int, int]:
    """(len_t, len_l) closest in token count to (1-mask_ratio)*T*L over
    [2..T] x [2..L]; always consumes one ``random.sample`` draw
    (/root/reference/Models.py:484-493)."""
    cands = [(t, l) for t in range(2, T + 1) for l in range(2, L + 1)]
    if not cands:
        raise ValueError(f"token grid {T}x{L} is too small for structured masking (needs >= 2x2)")
    target = (1 - mask_ratio) * T * L
    counts = torch.tensor([t * l for t, l in cands])
    gap = abs(target - counts)          # same float32 tensor arithmetic as the reference
    ties = torch.where(gap == torch.min(gap))[0]
    pick = random.sample(range(len(ties)), 1)[0]
    t, l = cands[int(ties[pick])]
    return int(t), int(l)


def _sincos_1d(width: int, pos: np.ndarray) -> np.ndarray:
    half = width // 2
    freq = np.arange(half, dtype=np.float32)
    freq /= width / 2.0
    freq = 1.0 / 10000 ** freq
    ang = np.einsum("m,d->md", pos.reshape(-1), freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_table(width: int, T: int, G: int) -> torch.Tensor:
    """Frozen 3-D sin-cos table [1, T*G*G, width]: first half encodes the spectral
    group, second half the spatial grid position (/root/reference/Models.py:11-47)."""
    if width % 4 != 0:
        raise ValueError("position-table width must be a multiple of 4")
    half = width // 2
    gw, gh = np.meshgrid(np.arange(G, dtype=np.float32), np.arange(G, dtype=np.float32))
    spatial = np.concatenate([_sincos_1d(half // 2, gw), _sincos_1d(half // 2, gh)], axis=1)
    spectral = _sincos_1d(half, np.arange(T, dtype=np.float32))
    table = np.concatenate([np.repeat(spectral[:, None, :], G * G, axis=1),
                            np.repeat(spatial[None, :, :], T, axis=0)], axis=-1)
    return torch.tensor(table.reshape(-1, width), dtype=torch.float32).unsqueeze(0)


def draw_drop_factors(rates_split: List[float], rates_fusion: List[float], n: int, len_t: int, len_l: int,
                      device, training: bool) -> List[Optional[torch.Tensor]]:
    """Stochastic-depth factors in the order the reference draws them
    (/root/reference/Models.py:244-251 via Block.forward :304-305 inside
    forward_encoder :879-891): all spatial blocks (attention branch then MLP
    branch), then all spectral blocks, then the fusion blocks.  "Per sample"
    means per row of the regrouped batch: n*len_t rows in the spatial encoder,
    n*len_l in the spectral one, n in the fusion blocks.  Entries are ``None``
    where the reference installs ``nn.Identity`` (rate 0) or when not training."""
    out: List[Optional[torch.Tensor]] = []

    def draw(rate: float, rows: int):
        if rate == 0.0 or not training:
            return None
        keep = 1.0 - rate
        f = torch.empty((rows, 1, 1), device=device, dtype=torch.float32).bernoulli_(keep)
        if keep > 0.0:
            f.div_(keep)
        return f

    for rows, rates in ((n * len_t, rates_split), (n * len_l, rates_split), (n, rates_fusion)):
        for r in rates:
            out.append(draw(r, rows))
            out.append(draw(r, rows))
    return out
