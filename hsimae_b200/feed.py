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
            s = np.asarray(s)
            if s.ndim != 3 or s.shape[2] != self.bands:
                raise ValueError("every scene must be [H, W, %d]" % self.bands)
            hw.append((s.shape[0], s.shape[1])); off.append(pos); pos += s.size
            flat.append(torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)).reshape(-1))
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
        self.scenes = torch.cat(flat).to(self.device) if flat else torch.empty(0, device=self.device)
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
        _lib.check(L.hsimae_gather_patches(self.scenes.data_ptr(), self.scene_off.data_ptr(), self.scene_hw.data_ptr(), self.bands, self.img,
                                           self.cut_info.data_ptr(), idx_d.data_ptr(), fl_d.data_ptr() if fl_d is not None else None,
                                           n, out.data_ptr(), st), "gather_patches")
        return out

    def epoch(self, batch_size: int, shuffle: bool = True):
        """Batches of one epoch in the order `DataLoader(dataset, batch_size, shuffle=True)` visits them (`:76`): the
        permutation comes from torch's RNG like RandomSampler's, the flips from Python's `random` sample by sample."""
        order = loader_order(self._len, shuffle)
        for i in range(0, self._len, batch_size):
            yield self.batch(order[i:i + batch_size])
