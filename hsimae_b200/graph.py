"""CUDA-graph execution of the latency-bound training steps.

The dual-branch fine-tuning step of the reference (`/root/reference/Model_Finetuning.py:147-166`: 32 labelled + 71
unlabelled tiles) is ~1000 kernels on 103 samples: the device work is a few hundred microseconds, the step is bound by
launch cost.  Every entry point of the C ABI only enqueues on the stream it is given and allocates nothing, so a whole
forward + backward can be captured once per (batch shape, visible shape) and replayed as ONE graph launch; the optimiser
(driver-owned, not capturable as `torch.optim.AdamW` is constructed by the drivers) runs after the replay on the static
gradient tensors.

Random numbers keep the reference's semantics: the visible shape is drawn with Python's `random` exactly as the eager
forward draws it (one `random.sample` per call, `Models.py:484-493`) and selects the graph; the mask noise and the
stochastic-depth factors are torch CUDA-generator draws inside the graph, which torch advances on every replay.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import modules as _mod
from .host import choose_visible_shape


class _Captured:
    __slots__ = ("graph", "x", "xu", "y", "loss", "logits", "grads")


class GraphedFinetuneStep:
    """`loss, logits = step(x, x_u, y)` == the body of the reference's fine-tuning loop

        loss_rec, _, _, outputs = model(x, x_u, mask_ratio=mask_ratio)
        loss = lamda * loss_rec + criterion(outputs, y)
        optimizer.zero_grad(); loss.backward(); optimizer.step()

    with forward + loss + backward replayed from a CUDA graph.  `model` is a `DualViT` in train mode on a CUDA device.
    The returned tensors are static buffers of the graph: read them before the next call."""

    def __init__(self, model, optimizer, criterion: Callable, lamda: float = 10.0, mask_ratio: float = 0.8, warmup: int = 2):
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("hsimae_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        self.model, self.opt, self.crit = model, optimizer, criterion
        self.lamda, self.mask_ratio, self.warmup = float(lamda), float(mask_ratio), int(warmup)
        self.graphs: Dict[Tuple, _Captured] = {}

    def _fwd_bwd(self, x, xu, y):
        loss_rec, _, _, logits = self.model(x, xu, mask_ratio=self.mask_ratio)
        loss = self.lamda * loss_rec + self.crit(logits, y)
        loss.backward()
        return loss, logits

    def _capture(self, shape, x, xu, y) -> _Captured:
        c = _Captured()
        c.x, c.xu, c.y = x.clone(), xu.clone(), y.clone()
        orig = _mod.choose_visible_shape
        _mod.choose_visible_shape = lambda T, L, r: shape      # the shape was drawn by __call__: do not draw again
        try:
            # eager warm-up on a side stream: first-use configuration of every kernel variant of this shape
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self.opt.zero_grad(set_to_none=True)
                    self._fwd_bwd(c.x, c.xu, c.y)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            # gradients must be (re)created inside the capture so that they live in the graph's memory pool
            self.opt.zero_grad(set_to_none=True)
            self.model._prepare(c.x)                            # the one-off (synchronising) upload of the packing table
            self.model.invalidate_weight_cache()                # the parameter re-pack itself is part of every replay
            c.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(c.graph):
                c.loss, c.logits = self._fwd_bwd(c.x, c.xu, c.y)
        finally:
            _mod.choose_visible_shape = orig
        # every graph writes its own static gradient tensors: a replay re-attaches them to the parameters
        c.grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]
        return c

    def __call__(self, x, xu, y):
        pe = self.model.patch_embed
        T, L = pe.b_grid_size, pe.grid_size ** 2
        shape = choose_visible_shape(T, L, self.mask_ratio)     # consumes the Python RNG exactly like the eager forward
        key = (shape, tuple(x.shape), tuple(xu.shape))
        c = self.graphs.get(key)
        if c is None:
            c = self.graphs[key] = self._capture(shape, x, xu, y)
        c.x.copy_(x, non_blocking=True); c.xu.copy_(xu, non_blocking=True); c.y.copy_(y, non_blocking=True)
        for p, g in c.grads:
            p.grad = g
        c.graph.replay()
        self.opt.step()
        return c.loss, c.logits
