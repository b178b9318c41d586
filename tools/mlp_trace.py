"""Per-role wait-cycle breakdown of the fused gated-MLP kernel (library built with HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops, _lib
L = C.CDLL(str(_lib.LIB_PATH))
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
NAMES = ["prod13: total - empty13", "prod2 : total empty2", "mma   : total full13 chunk_done out_empty full2 a_full",
         "gate  : total abfull tmem_ld math tmem_st fence+arrive", "final : total out_full", "prodA : total a_empty"]
for name, M, D, H in (("encoder", 4096 * 18, 256, 688), ("decoder", 4096 * 36, 64, 176)):
    x, w13, w2 = bf(M, D), bf(2 * H, D) * 0.05, bf(D, H) * 0.05
    b13, b2 = torch.zeros(2 * H, device=dev), torch.zeros(D, device=dev)
    resid = torch.randn(M, D, device=dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    for label, kw in (("train", dict(gamma=gamma, beta=beta)), ("noln,nog", dict(keep_g=False))):
        for _ in range(2): ops.mlp_fused(x, w13, b13, w2, b2, resid, **kw)
        torch.cuda.synchronize()
        buf = (C.c_longlong * (256 * 64))()
        L.hsimae_debug_trace_fused(buf, 256 * 64)
        t = torch.tensor(list(buf), dtype=torch.float64).view(256, 8, 8)[:148]
        print("==", name, label)
        for r in range(6):
            lead, peer = t[0::2, r].mean(0).tolist(), t[1::2, r].mean(0).tolist()
            print("  %-55s" % NAMES[r], " ".join("%8.0f" % v for v in lead[:6]), "| peer:", " ".join("%8.0f" % v for v in peer[:6]))
