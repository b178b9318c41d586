"""Fused optimiser step and LR schedule for the training loops -- opt-in replacements for the
`AdamW(...)` / `CosineLRScheduler(...)` objects the reference drivers build
(/root/reference/Model_Pretraining.py:80-88,100-104; Model_Finetuning.py:98-106,165-166,234).

`FusedAdamW` keeps `torch.optim.AdamW`'s constructor and `param_groups` / `zero_grad` / `step` surface; the update of
all parameters is ONE kernel launch over a device-resident job table (`hsimae_adamw_step`, csrc/optim.cu) instead of
~18 foreach launches over 535 tensors.  `CosineLRScheduler` restates timm 0.9.12's scheduler for the arguments the
drivers pass; timm is not available offline, so that formula is NOT pinned against timm itself."""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable

import numpy as np
import torch

from . import _lib

_JOB = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i4"), ("decay", "<f4"), ("tile0", "<i4"), ("pad", "<i4")])


class FusedAdamW(torch.optim.Optimizer):
    """AdamW with torch's semantics (decoupled weight decay, bias correction, eps outside the square root)."""

    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tile = None
        self._tables = {}     # (group index, step count slot) -> cached device job table; never part of state_dict()

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not isinstance(st["step"], int):
            st["step"] = int(st["step"])   # torch.optim.AdamW checkpoints keep the count as a 0-dim tensor
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _lib.load()
        if self._tile is None:
            self._tile = int(L.hsimae_adamw_tile_elems())
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda:
                    raise RuntimeError("hsimae_b200.optim.FusedAdamW updates CUDA parameters only")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
            lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
            decay = float(np.float32(1 - lr * wd))   # the Python scalar torch hands its foreach kernel, rounded once
            # bias correction depends on the per-parameter step count (a parameter that had no gradient for a while
            # lags behind, as in torch): one launch per distinct count -- one in practice
            by_step = {}
            for p in ps:
                by_step.setdefault(self._state(p)["step"], []).append(p)
            tables = self._tables
            for t0, sub in by_step.items():
                states = [self.state[p] for p in sub]
                # the table holds pointers only (the decay is a launch scalar): it survives every learning-rate change
                key = tuple((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                            for p, st in zip(sub, states))
                slot = (gi, len(sub), sub[0].data_ptr())
                cached = tables.get(slot)
                if cached is None or cached[0] != key:
                    jobs = np.zeros(len(sub), dtype=_JOB)
                    tile_job, tile0 = [], 0
                    for i, (p, st) in enumerate(zip(sub, states)):
                        nt = (p.numel() + self._tile - 1) // self._tile
                        jobs[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), 0.0, tile0, 0)
                        tile_job.append(np.full(nt, i, dtype=np.int32))
                        tile0 += nt
                    blob = np.concatenate([jobs.view(np.uint8), np.concatenate(tile_job).view(np.uint8)])
                    cached = (key, torch.from_numpy(blob).to(sub[0].device), len(sub), tile0)
                    tables[slot] = cached
                _, table, njobs, ntiles = cached
                t = t0 + 1
                bc1 = 1 - b1 ** t
                bc2_sqrt = math.sqrt(1 - b2 ** t)
                st_ptr = C.c_void_p(torch.cuda.current_stream(sub[0].device).cuda_stream)
                with torch.cuda.device(sub[0].device):          # the launch goes to the current device's context
                    _lib.check(L.hsimae_adamw_step(C.c_void_p(table.data_ptr()), njobs, ntiles, decay, 1 - b1, b2, 1 - b2, eps, bc2_sqrt,
                                                   -(lr / bc1), st_ptr), "adamw_step")
                for st in states:
                    st["step"] = t
            # the kernel wrote through raw pointers: tell torch (and the model's packed-weight cache, which watches
            # `_version`) that these tensors changed
            torch.autograd.graph.increment_version(ps)
        return loss


class CosineLRScheduler:
    """timm's cosine schedule with linear warm-up, for `CosineLRScheduler(optimizer, t_initial=iters, lr_min=1e-6,
    warmup_t=...)` followed by `scheduler.step(iter_num)` (`Model_Pretraining.py:88,104`): the optimiser starts at
    `warmup_lr_init`; `step(t)` sets the rate for index t: linear from `warmup_lr_init` to the base rate over
    `warmup_t` steps, then `lr_min + (base - lr_min)(1 + cos(pi (t - warmup_t) / (t_initial - warmup_t ... )))/2`
    (timm's `warmup_prefix=False`: the cosine clock is t itself).  Restated from memory of timm 0.9.12 -- unpinned."""

    def __init__(self, optimizer, t_initial: int, lr_min: float = 0.0, warmup_t: int = 0, warmup_lr_init: float = 0.0):
        self.optimizer = optimizer
        self.t_initial, self.lr_min, self.warmup_t, self.warmup_lr_init = t_initial, lr_min, warmup_t, warmup_lr_init
        self.base_values = [g["lr"] for g in optimizer.param_groups]
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self._set([warmup_lr_init for _ in self.base_values])
        else:
            self.warmup_steps = [1 for _ in self.base_values]

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def get_lr(self, t: int):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        if t < self.t_initial:
            return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t / self.t_initial)) for v in self.base_values]
        return [self.lr_min for _ in self.base_values]

    def step(self, t: int):
        self._set(self.get_lr(t))
