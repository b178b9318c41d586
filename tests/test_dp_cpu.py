"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (hsimae_b200/dp.py) -- bucket order, averaging,
parameter broadcast.  The device work is not exercised here (no GPU); the same GradSync object drives NCCL on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import TINY


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import Models as M
        from hsimae_b200 import dp
        torch.manual_seed(100 + rank)          # replicas start different ...
        model = M.HSIMAE(**TINY)
        dp.broadcast_parameters(model)          # ... and are made identical
        ref = [p.detach().clone() for p in model.parameters()]
        gathered = [torch.zeros_like(ref[5]) for _ in range(world)]
        dist.all_gather(gathered, ref[5])
        assert all(torch.equal(g, gathered[0]) for g in gathered)

        sync = dp.attach(model)
        rt = model._runtime()
        assert sum(n for _, n in sync.buckets) == rt.grad_elems and len(sync.buckets) == 6
        grads = torch.arange(rt.grad_elems, dtype=torch.float32) * (rank + 1)
        step = sync.begin(grads)
        for b in (5, 4, 3, 2, 1):               # backward-completion order; bucket 0 is flushed by finish()
            step.ready(b)
        step.ready(5)                           # idempotent
        step.finish()
        expect = torch.arange(rt.grad_elems, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        assert torch.allclose(grads, expect)
        # decoder bucket is the tail of the arena, patch-embed + the first third of the spatial encoder the head
        off5, n5 = sync.buckets[5]
        assert off5 + n5 == rt.grad_elems and sync.buckets[0][0] == 0
        named = dict(zip(rt.names, rt.grad_off))
        assert named["decoder_pred.weight"] >= off5 and named["patch_embed.proj.weight"] < sync.buckets[0][1]
        sd = TINY["s_depth"]
        assert named["blocks_1.0.attn.q.weight"] < sync.buckets[1][0] or sd < 3
        assert sync.buckets[2][0] <= named[f"blocks_1.{sd - 1}.attn.q.weight"] < sync.buckets[3][0]
        assert sync.buckets[3][0] <= named["blocks_2.0.attn.q.weight"] < sync.buckets[4][0] <= named["blocks.0.attn.q.weight"] < off5
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_attach_requires_process_group():
    import Models as M
    from hsimae_b200 import dp
    with pytest.raises(RuntimeError):
        dp.attach(M.HSIMAE(**TINY))
