"""Fused dgrad + LayerNorm-backward kernel (kEpiLnBwd) at the pretraining shapes: time, and (library built with
HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE) the per-role wait-cycle breakdown."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops, _lib
L = _lib.load()
dev = "cuda"; B = 4096; M, D, H = B * 18, 256, 688
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x = torch.randn(M, D, device=dev); stats = torch.stack([x.mean(1), (x.var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
gamma = torch.ones(D, device=dev); dx = torch.randn(M, D, device=dev)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
trace = hasattr(C.CDLL(str(_lib.LIB_PATH)), "hsimae_debug_trace") and os.environ.get("TRACE")
for K in (1376, 768, 256):
    A, W = bf(M, K), bf(D, K) * 0.05
    fused = t(lambda: ops.gemm_lnbwd(A, W, x, stats, gamma, dx, dgamma=dg, dbeta=db, inplace=True))
    plain = t(lambda: ops.gemm(A, W, 0))
    print("K=%d fused %.1f us | dgrad alone %.1f us" % (K, fused, plain))
    if trace:
        buf = (C.c_longlong * (256 * 16))()
        ops.gemm_lnbwd(A, W, x, stats, gamma, dx, dgamma=dg, dbeta=db, inplace=True); torch.cuda.synchronize()
        C.CDLL(str(_lib.LIB_PATH)).hsimae_debug_trace(buf, 256 * 16)
        tt = torch.tensor(list(buf), dtype=torch.float64).view(256, 4, 4)[:148]
        names = ["producer: total, empty-wait", "mma: total, -, tempty-wait, full-wait", "epilogue w2: total, tfull-wait, in-epilogue(incl tfull)"]
        for r in range(3):
            print("  %-60s" % names[r], " ".join("%9.0f" % v for v in tt[:, r].mean(0).tolist()))
