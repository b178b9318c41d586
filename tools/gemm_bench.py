import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops
dev = "cuda"; B = 4096; M, D, H = B * 18, 256, 688
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x, w13, wqkv, w2 = bf(M, D), bf(2 * H, D) * 0.05, bf(3 * D, D) * 0.05, bf(D, H) * 0.05
g, resid = bf(M, H), torch.randn(M, D, device=dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dab = bf(M, 2 * H); w2t = bf(H, D) * 0.05; w13t = bf(D, 2 * H) * 0.05; wp = bf(D, D) * 0.05
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
print(os.environ.get("HSIMAE_GEMM_STAGES", "-"),
      "qkv %.1f" % t(lambda: ops.gemm(x, wqkv, 0)),
      "swiglu %.1f" % t(lambda: ops.gemm(x, w13, 3)),
      "w2+ln %.1f" % t(lambda: ops.gemm(g, w2, 2, resid=resid, gamma=gamma, beta=beta)),
      "proj+ln %.1f" % t(lambda: ops.gemm(x, wp, 2, resid=resid, gamma=gamma, beta=beta)),
      "dswiglu %.1f" % t(lambda: ops.gemm(x, w2t, 4, ab=dab)),
      "swiglu(no ab) %.1f" % t(lambda: ops.gemm(x, w13, 3, keep_ab=False)),
      "dgate(recompute) %.1f" % t(lambda: ops.gemm(x, w2t, 5, A2=x, B2=w13)),
      "dgrad1376 %.1f" % t(lambda: ops.gemm(dab, w13t, 0)),
      "dgrad256 %.1f" % t(lambda: ops.gemm(x, wp, 0)))

gw1, gw3 = torch.zeros(684, D, device=dev), torch.zeros(684, D, device=dev)
gb1, gb3 = torch.zeros(684, device=dev), torch.zeros(684, device=dev)
gw2, gb2 = torch.zeros(D, 684, device=dev), torch.zeros(D, device=dev)
dqkv = bf(M, 3 * D); gq, gqb = torch.zeros(3 * D, D, device=dev), torch.zeros(3 * D, device=dev)
gp, gpb = torch.zeros(D, D, device=dev), torch.zeros(D, device=dev)
print("wgrad pair=%s -" % os.environ.get("HSIMAE_WGRAD_PAIR", "1"),
      "dW13 %.1f" % t(lambda: ops.wgrad(dab, x, gw1, dst1=gw3, row_map=1, rows_valid=684, bias0=gb1, bias1=gb3)),
      "dW2 %.1f" % t(lambda: ops.wgrad(x, g, gw2, cols_valid=684, bias0=gb2)),
      "dWqkv %.1f" % t(lambda: ops.wgrad(dqkv, x, gq, bias0=gqb)),
      "dWproj %.1f" % t(lambda: ops.wgrad(x, x, gp, bias0=gpb)))
