"""On-device pretraining data feed: the scenes live in HBM, one kernel launch assembles a batch.

Host-side mirror of the reference's `HSIdataset4PT` (/root/reference/Model_Pretraining.py:21-54): same constructor
arguments (`data_cubes = [list of HWC scenes, int16 cut_info]`, `train`), same per-sample semantics and the same
consumption of Python's `random` for the flips, so a seeded run sees the batches the reference's DataLoader would have
produced -- without the per-sample Python slicing, the host-side collation and the H2D copy."""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def loader_order(n: int, shuffle: bool = True) -> torch.Tensor:
    """Sample order of one pass over `DataLoader(dataset, shuffle=shuffle, num_workers=0)` (`Model_Pretraining.py:76`),
    consuming torch's global RNG the way the loader does: one int64 draw for the iterator's base seed, then (when
    shuffling) one for RandomSampler's private generator, whose `randperm(n)` is the order."""
    torch.empty((), dtype=torch.int64).random_()           # _BaseDataLoaderIter: base seed
    if not shuffle:
        return torch.arange(n)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def initial_seq(length: int, size: int, stride: int) -> np.ndarray:
    """Window origins along one axis (`get_inital_seq`, Utils/Preprocessing.py:8-20): step `size // stride`, the last
    origin moved back so that the final window ends at the border."""
    n1 = length // size
    l_r = length - n1 * size
    step = int(size // stride)
    n2 = l_r // step
    num = int((n1 - 1) * stride + n2 + (1 if l_r - n2 * step == 0 else 2))
    seq = np.arange(0, num * step, step)
    seq[-1] = length - size
    return seq


def split_info(shape, target_size, stride, num, max_, min_) -> list:
    """Rows `(c, h, w, scene, max, min)` of one scene's cut table (`get_split_info`, Utils/Preprocessing.py:67-79):
    the cartesian product channel-origin x row-origin x column-origin, in that nesting order."""
    from itertools import product
    rows, cols, ch = shape
    ch_seq = initial_seq(ch, target_size[2], stride[2])
    row_seq = initial_seq(rows, target_size[0], stride[0])
    col_seq = initial_seq(cols, target_size[1], stride[1])
    return list(product(ch_seq, row_seq, col_seq, [num], [max_], [min_]))


def get_data_cut_file(data_path, patch_size=9, save_path=None, norm=False, GWPCA=True, ratio=1.0, device="cuda:0", sign="auto"):
    """Mirror of `get_data_cut_file` (Utils/Preprocessing.py:82-118): load every `[h, w, bands]` .npy scene, reduce it with
    the group-wise PCA -- here on the device (hsimae_b200.gwpca) -- and build the int16 cut table of 9 x 9 windows.
    Returns `[data_cubes, cut_locs]` like the reference; `data_cubes` are device tensors when `GWPCA` (numpy arrays
    otherwise), ready for :class:`PatchFeed`.  Reference quirks kept: `patch_size` is ignored (windows are 9 x 9 x all
    channels, `:101,104`); the first 14 scenes use origin step 3, are row-shuffled with numpy's global RNG and cut to
    `ratio`, later scenes use non-overlapping windows and neither (`:100-110`); max / min are truncated by the int16 cast."""
    from .gwpca import applyGWPCA
    data_cubes, cut_locs = [], []
    for num_count, path in enumerate(data_path):
        scene = np.load(path) if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__") else path
        if GWPCA:
            scene = applyGWPCA(scene, nc=32, group=4, whiten=True, sign=sign, device=device)
        h, w, c = scene.shape
        if norm:
            max_, min_ = float(scene.max()), float(scene.min())
        else:
            max_, min_ = 1, 0
        if num_count >= 14:
            cut_loc = split_info((h, w, c), (9, 9, c), (1, 1, 1), num_count, max_, min_)
        else:
            cut_loc = np.array(split_info((h, w, c), (9, 9, c), (3, 3, 1), num_count, max_, min_))
            np.random.shuffle(cut_loc)
            cut_loc = list(cut_loc[:int(cut_loc.shape[0] * ratio)])
        cut_locs += cut_loc
        data_cubes.append(scene)
    cut_locs = np.array(cut_locs, dtype=np.int16)
    if save_path:
        np.save(save_path, cut_locs)
    return [data_cubes, cut_locs]


class PatchFeed:
    def __init__(self, data_cubes, train: bool = False, device="cuda:0", img: int = 9):
        scenes, cut_info = data_cubes[0], np.asarray(data_cubes[1])
        if cut_info.ndim != 2 or cut_info.shape[1] != 6:
            raise ValueError("cut_info must be [n, 6] rows of (c, h, w, scene, max, min)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("hsimae_b200.feed.PatchFeed assembles batches on a CUDA device only")
        self.train, self.img = train, img
        self.bands = int(scenes[0].shape[2])
        hw, off, flat = [], [], []
        pos = 0
        for s in scenes:
            if not isinstance(s, torch.Tensor):                     # device tensors (hsimae_b200.gwpca output) stay on the device
                s = torch.from_numpy(np.ascontiguousarray(s))
            if s.dim() != 3 or s.shape[2] != self.bands:
                raise ValueError("every scene must be [H, W, %d]" % self.bands)
            hw.append((s.shape[0], s.shape[1])); off.append(pos); pos += s.numel()
            flat.append(s.to(self.device, torch.float32).reshape(-1))
        ci = cut_info.astype(np.int16)
        # windows must lie inside their scene (the reference would silently return a short slice and fail in collation)
        for k, (h, w) in enumerate(hw):
            rows = ci[ci[:, 3] == k]
            if len(rows) and (rows[:, 1].min() < 0 or rows[:, 2].min() < 0 or rows[:, 1].max() + img > h or rows[:, 2].max() + img > w):
                raise ValueError(f"cut_info has windows outside scene {k} ({h} x {w})")
        if len(ci) and (ci[:, 3].min() < 0 or ci[:, 3].max() >= len(scenes)):
            raise ValueError("cut_info refers to a scene that does not exist")
        if len(ci) and np.any(ci[:, 4] == ci[:, 5]):
            raise ValueError("cut_info has max == min (division by zero in the reference as well)")
        self.scenes = torch.cat(flat) if flat else torch.empty(0, device=self.device)
        self.scene_off = torch.tensor(off, dtype=torch.int64, device=self.device)
        self.scene_hw = torch.tensor(hw, dtype=torch.int32, device=self.device).reshape(-1)
        self.cut_info = torch.from_numpy(ci).to(self.device)
        self._len = len(ci)

    def __len__(self):
        return self._len

    def draw_flips(self, n: int) -> Optional[torch.Tensor]:
        """(hflip, vflip) per sample, drawn from Python's `random` in the reference's order (`Model_Pretraining.py:28-38,46-48`)."""
        if not self.train:
            return None
        f = np.zeros((n, 2), dtype=np.uint8)
        for i in range(n):
            f[i, 0] = random.random() < 0.5
            f[i, 1] = random.random() < 0.5
        return torch.from_numpy(f)

    def batch(self, indices: Sequence[int], flips: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, 1, bands, img, img] fp32 on the device == default_collate([dataset[i] for i in indices])."""
        idx = torch.as_tensor(indices, dtype=torch.int64)
        n = int(idx.numel())
        if n and (int(idx.min()) < 0 or int(idx.max()) >= self._len):
            raise IndexError("sample index out of range")
        if flips is None:
            flips = self.draw_flips(n)
        out = torch.empty(n, 1, self.bands, self.img, self.img, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        idx_d = idx.to(self.device, non_blocking=True)
        fl_d = flips.to(self.device, dtype=torch.uint8, non_blocking=True).contiguous() if flips is not None else None
        L = _lib.load()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):        # the launch goes to the current device's context
            _lib.check(L.hsimae_gather_patches(self.scenes.data_ptr(), self.scene_off.data_ptr(), self.scene_hw.data_ptr(), self.bands,
                                               self.img, self.cut_info.data_ptr(), idx_d.data_ptr(),
                                               fl_d.data_ptr() if fl_d is not None else None, n, out.data_ptr(), st), "gather_patches")
        return out

    def epoch(self, batch_size: int, shuffle: bool = True):
        """Batches of one epoch in the order `DataLoader(dataset, batch_size, shuffle=True)` visits them (`:76`): the
        permutation comes from torch's RNG like RandomSampler's, the flips from Python's `random` sample by sample."""
        order = loader_order(self._len, shuffle)
        for i in range(0, self._len, batch_size):
            yield self.batch(order[i:i + batch_size])
