"""Timings of the attention kernels at the bench shapes (Large, batch 4096)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops
B = 4096
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
res = []
for name, D, H, K, spec in [("spatial", 256, 16, 18, (2, 9, 9, 1)), ("spectral", 256, 16, 18, (9, 2, 1, 9)), ("spectral len3", 256, 16, 18, (6, 3, 1, 6)), ("spatial len6", 256, 16, 18, (3, 6, 6, 1)), ("fusion", 256, 16, 18, (1, 18, 18, 1)),
                            ("decoder", 64, 8, 36, (1, 36, 36, 1))]:
    qkv, do = bf(B * K, 3 * D), bf(B * K, D)
    out, lse = ops.attention_forward(qkv, B, D, H, K, *spec)
    f = t(lambda: ops.attention_forward(qkv, B, D, H, K, *spec))
    b = t(lambda: ops.attention_backward(qkv, out, lse, do, B, D, H, K, *spec))
    res.append("%s fwd %.1f bwd %.1f" % (name, f, b))
print(" | ".join(res))
