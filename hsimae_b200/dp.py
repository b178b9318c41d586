"""Data-parallel gradient exchange for the pretraining path.

One process per GPU; every rank holds a full replica and its own shard of the
batch.  The only exchange step is the mean of the parameter gradients.  The
native backward produces all gradients in ONE fp32 arena whose regions become
final in a known order (decoder -> fusion blocks + final norm -> spectral encoder
-> spatial encoder in thirds, the first third with the patch embedding, see hsimae_plan_grad_bucket), so each
region is all-reduced on NCCL's stream as soon as the stage that finishes it
has been enqueued, overlapping with the remaining backward kernels.  The
reference has no counterpart (single device, Model_Pretraining.py:59).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    """Attach to a model with :func:`attach`; used by the autograd node in modules.py."""

    def __init__(self, buckets, process_group=None):
        self.group = process_group
        self.buckets = list(buckets)   # [(offset, elems)] indexed by bucket id
        self.world = dist.get_world_size(process_group)
        self._comm_stream = None

    def begin(self, grads: torch.Tensor) -> "_Step":
        if grads.is_cuda and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=grads.device)
        return _Step(self, grads)


class _Step:
    def __init__(self, owner: GradSync, grads: torch.Tensor):
        self.o, self.grads = owner, grads
        self.works: List = []
        self.done = set()

    def ready(self, bucket: int) -> None:
        """the gradients of `bucket` are final on the current stream: start their all-reduce"""
        if bucket in self.done:
            return
        self.done.add(bucket)
        off, n = self.o.buckets[bucket]
        if n == 0 or self.o.world == 1:
            return
        view = self.grads[off:off + n]
        if view.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.o._comm_stream):
                self.o._comm_stream.wait_event(ev)
                self.works.append(dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.o.group, async_op=True))
        else:  # gloo (CPU tests): no AVG, no streams
            w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.o.group, async_op=True)
            self.works.append((w, view))

    def finish(self) -> None:
        """make the current stream wait for every outstanding exchange"""
        for b in range(len(self.o.buckets)):
            self.ready(b)
        for w in self.works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.o.world)
            else:
                w.wait()
        self.works.clear()


def attach(model, process_group=None) -> GradSync:
    """Enable gradient averaging across `process_group` for `model` (HSIMAE / DualViT / HSIViT)."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rt = model._runtime()
    sync = GradSync(rt.buckets, process_group)
    model.__dict__["_dp"] = sync
    return sync


def broadcast_parameters(model, src: int = 0, process_group=None) -> None:
    """make every replica start from rank `src`'s parameters"""
    with torch.no_grad():
        for p in model.parameters():
            dist.broadcast(p, src=src, group=process_group)   # in place on `p` itself: bumps `_version`
    # belt and braces: the packed bf16 / fp32 operand arenas are rebuilt on the next forward whatever torch recorded
    if hasattr(model, "invalidate_weight_cache"):
        model.invalidate_weight_cache()
