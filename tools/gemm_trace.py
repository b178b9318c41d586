"""Per-role wait-cycle breakdown of the A-resident GEMM (library built with HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops, _lib
L = _lib.load()
dev = "cuda"; B = 4096; M, D, H = B * 18, 256, 688
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x, w13, wqkv = bf(M, D), bf(2 * H, D) * 0.05, bf(3 * D, D) * 0.05
dab = bf(M, 2 * H); w2t = bf(H, D) * 0.05
def trace(name, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (256 * 16))()
    L.hsimae_debug_trace(buf, 256 * 16)
    t = torch.tensor(list(buf), dtype=torch.float64).view(256, 4, 4)[:148]
    lead = t[0::2] if os.environ.get("HSIMAE_GEMM_PAIR", "1") != "0" else t
    names = ["producer: total, empty-wait", "mma: total, a_full-wait, tempty-wait, full-wait", "epilogue w2: total, tfull-wait, in-epilogue(incl tfull)"]
    print("==", name)
    for r in range(3):
        print("  %-55s" % names[r], " ".join("%9.0f" % v for v in lead[:, r].mean(0).tolist()), "| all CTAs:", " ".join("%9.0f" % v for v in t[:, r].mean(0).tolist()))
g, resid = bf(M, H), torch.randn(M, D, device=dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
w2 = bf(D, H) * 0.05; w13t = bf(D, 2 * H) * 0.05; wp = bf(D, D) * 0.05
trace("w2+ln (generic, K=688, N=256)", lambda: ops.gemm(g, w2, 2, resid=resid, gamma=gamma, beta=beta))
trace("proj+ln (generic, K=256, N=256)", lambda: ops.gemm(x, wp, 2, resid=resid, gamma=gamma, beta=beta))
trace("dgrad1376 (generic, K=1376, N=256)", lambda: ops.gemm(dab, w13t, 0))
trace("dgrad256 (generic, K=256, N=256)", lambda: ops.gemm(x, wp, 0))
trace("qkv", lambda: ops.gemm(x, wqkv, 0))
trace("swiglu", lambda: ops.gemm(x, w13, 3))
trace("dswiglu", lambda: ops.gemm(x, w2t, 4, ab=dab))
trace("swiglu (no ab)", lambda: ops.gemm(x, w13, 3, keep_ab=False))
trace("dgate (recompute): epilogue row = total, tfull-wait", lambda: ops.gemm(x, w2t, 5, A2=x, B2=w13))
