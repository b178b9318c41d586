"""Fused gated-MLP kernel vs the two launches it replaces, at the bench shapes (tuning instrument)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops

dev = "cuda"
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters

bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
for name, M, D, H in (("encoder", 4096 * 18, 256, 688), ("decoder", 4096 * 36, 64, 176)):
    x, w13, w2 = bf(M, D), bf(2 * H, D) * 0.05, bf(D, H) * 0.05
    b13, b2 = torch.zeros(2 * H, device=dev), torch.zeros(D, device=dev)
    resid = torch.randn(M, D, device=dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    fl = 2 * M * 3 * H * D
    def unfused():
        u = ops.gemm(x, w13, ops.EPI_SWIGLU, bias=b13, keep_ab=False)
        ops.gemm(u["g"], w2, ops.EPI_RESID_LN, bias=b2, resid=resid, gamma=gamma, beta=beta)
    r = {"shape": name, "M": M,
         "fused_train_us": timeit(lambda: ops.mlp_fused(x, w13, b13, w2, b2, resid, gamma=gamma, beta=beta)),
         "fused_infer_us": timeit(lambda: ops.mlp_fused(x, w13, b13, w2, b2, resid, gamma=gamma, beta=beta, keep_g=False)),
         "fused_noln_us": timeit(lambda: ops.mlp_fused(x, w13, b13, w2, b2, resid, keep_g=False)),
         "unfused_us": timeit(unfused)}
    r["fused_train_tflops"] = fl / r["fused_train_us"] / 1e6
    print(json.dumps(r))
