"""Throughput of the on-device data feed vs the reference's host path: Salinas-shaped scene, batch 4096."""
import sys, os, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hsimae_b200.feed import PatchFeed
from oracle import feed_oracle as FO
rng = np.random.default_rng(0)
scene = rng.standard_normal((512, 217, 32)).astype(np.float32)
cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 504) for w in range(0, 209)], dtype=np.int16)
feed = PatchFeed([[scene], cut], train=True)
B = 4096
idx = torch.randint(0, len(cut), (B,))
flips = torch.randint(0, 2, (B, 2), dtype=torch.uint8)
idx_d, fl = idx, flips
for _ in range(3): feed.batch(idx, flips)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): x = feed.batch(idx, flips)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
byt = 2 * B * 32 * 81 * 4
t0 = time.perf_counter(); ref = FO.get_batch([scene], cut, idx[:512].numpy(), flips[:512].numpy()); t1 = time.perf_counter()
assert np.array_equal(x[:512].cpu().numpy(), ref)
print("device feed: %.3f ms per batch of %d (%.0f GB/s of read+write, %.1f M patches/s; includes the index/flip H2D copies) | "
      "host path (numpy restatement of the reference's per-sample loop, 1 core): %.1f ms per 512 = %.0f patches/s"
      % (ms, B, byt / ms / 1e6, B / ms / 1e3, (t1 - t0) * 1e3, 512 / (t1 - t0)))
