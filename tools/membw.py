"""Calibration: achievable HBM bandwidth for read-only, write-only and copy streams (torch kernels)."""
import torch
dev = "cuda"
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
def t(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(it):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return best * 1e-3
af = a.view(torch.float32)
print("read-only  (sum fp32)  %.0f GB/s" % (af.numel() * 4 / t(lambda: af.sum()) / 1e9))
print("write-only (fill)      %.0f GB/s" % (n * 2 / t(lambda: b.fill_(1.0)) / 1e9))
print("copy (r+w bytes)       %.0f GB/s" % (2 * n * 2 / t(lambda: b.copy_(a)) / 1e9))
