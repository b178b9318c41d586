import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops
M, D = 4096 * 18, 256
x = torch.randn(M, D, device="cuda").to(torch.bfloat16); w = (torch.randn(D, D, device="cuda") * 0.05).to(torch.bfloat16)
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) * 1e3 / n
o = ops.gemm(x, w, ops.EPI_BIAS_BF16)["out"].float()
ref = x.float() @ w.float().t()
print(json.dumps({"N256": os.environ.get("HSIMAE_GEMM_N256", "1"), "us": t(lambda: ops.gemm(x, w, ops.EPI_BIAS_BF16)), "err": float((o - ref).norm() / ref.norm())}))
