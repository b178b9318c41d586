"""GPU, >= 2 devices: two data-parallel ranks (NCCL) produce the same averaged gradients as one rank on the
concatenated batch.  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, random, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HS_ROOT"])
import Models, hsimae_b200.modules as mod
from hsimae_b200 import dp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=128, depth=12, num_heads=8, s_depth=9,
           decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
torch.manual_seed(3); random.seed(3)
model = Models.HSIMAE(**cfg).to(dev)
dp.broadcast_parameters(model)
g = torch.Generator(device="cpu").manual_seed(11)
x_all = torch.randn(64, 1, 32, 9, 9, generator=g).to(dev)
nt_all, nl_all = torch.rand(64, 4, generator=g).to(dev), torch.rand(64, 9, generator=g).to(dev)
def run(x, nt, nl):
    feed = [nt.contiguous(), nl.contiguous()]
    orig_s, orig_r = mod.choose_visible_shape, torch.rand
    mod.choose_visible_shape = lambda T, L, r: (3, 6); torch.rand = lambda *a, **k: feed.pop(0)
    try:
        model.zero_grad(); loss, _, _ = model(x, mask_ratio=0.5)
    finally:
        mod.choose_visible_shape, torch.rand = orig_s, orig_r
    loss.backward()
    return loss.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
full_loss, full = run(x_all, nt_all, nl_all)            # single-replica reference on the global batch
dp.attach(model)
sl = slice(rank * 32, rank * 32 + 32)
loss, part = run(x_all[sl], nt_all[sl], nl_all[sl])     # sharded step with gradient averaging
worst = max(float((part[k] - full[k]).norm() / (full[k].norm() + 1e-12)) for k in full if not k.endswith("attn.k.bias"))
lt = torch.tensor([loss], device=dev); dist.all_reduce(lt); 
ok = worst < 5e-3 and abs(lt.item() / world - full_loss) < 1e-4
print(f"rank {rank} worst grad rel err {worst:.2e} mean loss {lt.item()/world:.6f} vs {full_loss:.6f} -> {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_ranks_match_single_rank(tmp_path):
    script = tmp_path / "dp_check.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, HS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2, r.stdout
