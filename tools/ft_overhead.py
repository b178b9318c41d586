import json, os, sys, time, random
sys.path.insert(0, "/root/repo")
import torch
import Models
torch.manual_seed(0); random.seed(0)
model = Models.DualViT(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=256, depth=12, num_heads=16,
                       s_depth=9, drop_path=0.2, decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True,
                       trunc_init=True).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-2)
crit = torch.nn.CrossEntropyLoss(ignore_index=0)
nl, nu = 32, 71
x, xu = torch.randn(nl, 1, 32, 9, 9, device="cuda"), torch.randn(nu, 1, 32, 9, 9, device="cuda")
y = torch.randint(1, 17, (nl,), device="cuda")
def step():
    loss_rec, _, _, logits = model(x, xu, mask_ratio=0.8)
    loss = 10 * loss_rec + crit(logits, y)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
# GPU time via events, CPU enqueue time via perf_counter without sync
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(20): step()
b.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("cpu enqueue %.2f ms/step, gpu span %.2f ms/step, wall %.2f" % ((t1 - t0) / 20 * 1e3, a.elapsed_time(b) / 20, (t2 - t0) / 20 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev) / 3 / 1e3
print("sum of kernel durations %.2f ms/step over %d launches/step" % (tot, sum(e.count for e in ev) / 3))
rows = sorted(ev, key=lambda e: -e.device_time_total)[:18]
for e in rows:
    print("%8.1f us/step  n=%5.1f  avg %6.1f us  %s" % (e.device_time_total / 3, e.count / 3, e.device_time_total / max(e.count, 1), e.key[:90]))
