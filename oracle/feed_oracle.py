"""CPU restatement of the reference's pretraining data feed -- TEST INFRASTRUCTURE ONLY (see oracle/hsimae_oracle.py).

Follows `HSIdataset4PT.__getitem__` (/root/reference/Model_Pretraining.py:40-51) and the default collation of the
`DataLoader(train_dataset, batch_size=bs, shuffle=True, num_workers=0)` at `:76`.  Pinned against the reference class
itself (imported with a stub for the absent `timm`) in tests/test_feed_cpu.py and against tests/golden/feed.npz
(oracle/make_golden_feed.py)."""
from __future__ import annotations

import random

import numpy as np


def draw_flips(n: int, train: bool = True) -> np.ndarray:
    """Flip decisions for `n` consecutive samples, consuming Python's `random` exactly as the reference does:
    per sample one draw for the horizontal flip, then one for the vertical flip (`Model_Pretraining.py:28-38,46-48`);
    none when `train` is False.  Returns uint8 [n, 2] = (hflip, vflip)."""
    f = np.zeros((n, 2), dtype=np.uint8)
    if train:
        for i in range(n):
            f[i, 0] = random.random() < 0.5
            f[i, 1] = random.random() < 0.5
    return f


def get_item(data_cubes, cut_info, index: int, flips=(0, 0), img: int = 9) -> np.ndarray:
    """One sample, [1, C, img, img] float32 (`Model_Pretraining.py:40-51`)."""
    c, h, w, num, max_, min_ = cut_info[index]                 # :41  (int16 row; `c` is not used by the reference either)
    cube = data_cubes[num]                                     # :42
    data = cube[h:h + img, w:w + img, :]                       # :43
    data = (data - min_) / (max_ - min_)                       # :44
    if flips[0]:
        data = np.flip(data, 1)                                # :30  horizontal: the W axis of [H, W, C]
    if flips[1]:
        data = np.flip(data, 0)                                # :36  vertical: the H axis
    data = np.ascontiguousarray(data, dtype=np.float32)        # :49
    return np.transpose(data[None], (0, 3, 1, 2))              # :50  [1, H, W, C] -> [1, C, H, W]


def get_batch(data_cubes, cut_info, indices, flips=None, img: int = 9) -> np.ndarray:
    """Default-collated batch, [B, 1, C, img, img] float32."""
    if flips is None:
        flips = np.zeros((len(indices), 2), dtype=np.uint8)
    if len(indices) == 0:
        bands = data_cubes[0].shape[2]
        return np.zeros((0, 1, bands, img, img), dtype=np.float32)
    return np.stack([get_item(data_cubes, cut_info, int(i), flips[k], img) for k, i in enumerate(indices)], axis=0)
