"""Launch each hot kernel once at the bench shapes (M = 4096*18 encoder rows) so ncu can capture them individually."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops

dev = "cuda"
B = 4096
M, D, H = B * 18, 256, 688
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x, w13, wqkv, w2, wp = bf(M, D), bf(2 * H, D) * 0.05, bf(3 * D, D) * 0.05, bf(D, H) * 0.05, bf(D, D) * 0.05
g, resid = bf(M, H), torch.randn(M, D, device=dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dab, dqkv, qkv = bf(M, 2 * H), bf(M, 3 * D), bf(M, 3 * D)
gw1, gw3, gb1, gb3 = (torch.zeros(684, D, device=dev) for _ in range(2)) , None, None, None
gw1, gw3 = torch.zeros(684, D, device=dev), torch.zeros(684, D, device=dev)
gb1, gb3 = torch.zeros(684, device=dev), torch.zeros(684, device=dev)
w2t = bf(H, D) * 0.05
x2 = bf(M, D)
b13, b2 = torch.zeros(2 * H, device=dev), torch.zeros(D, device=dev)
w13t = bf(D, 2 * H) * 0.05
wqkvt = bf(D, 3 * D) * 0.05
xf, dxf = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
stats = torch.stack([xf.mean(1), (xf.var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
dgam, dbet = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
z = lambda *s_: torch.zeros(*s_, device=dev)
wg_jobs = [dict(Y=x, X=g, dst0=z(D, 684), cols_valid=684, bias0=z(D)),
           dict(Y=dab, X=x2, dst0=gw1, dst1=gw3, row_map=1, rows_valid=684, bias0=gb1, bias1=gb3),
           dict(Y=x, X=x2, dst0=z(D, D), bias0=z(D)),
           dict(Y=dqkv, X=x2, dst0=z(3 * D, D), bias0=z(3 * D))]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    torch.cuda.profiler.start()
    ops.gemm(x, wqkv, ops.EPI_BIAS_BF16)                                      # qkv projection
    ops.gemm(x, w13, ops.EPI_SWIGLU, keep_ab=False)                           # gated up-projection (training path: g only)
    ops.gemm(g, w2, ops.EPI_RESID_LN, resid=resid, gamma=gamma, beta=beta)    # down-projection + residual + LN
    ops.gemm(x, w2t, ops.EPI_DGATE, A2=x2, B2=w13)                            # d(gate), a|b recomputed (two distinct activations)
    ops.gemm(dab, w13t, ops.EPI_BIAS_BF16)                                    # dgrad K=1376
    ops.wgrad(dab, x, gw1, dst1=gw3, row_map=1, rows_valid=684, bias0=gb1, bias1=gb3)   # dW13
    ops.wgrad(x, g, torch.zeros(D, 684, device=dev), cols_valid=684, bias0=torch.zeros(D, device=dev))  # dW2
    out, lse = ops.attention_forward(qkv, B, D, 16, 18, 1, 18, 18, 1)        # fusion attention
    ops.attention_backward(qkv, out, lse, x, B, D, 16, 18, 1, 18, 18, 1)
    out, lse = ops.attention_forward(qkv, B, D, 16, 18, 3, 6, 6, 1)          # spatial
    out, lse = ops.attention_forward(qkv, B, D, 16, 18, 6, 3, 1, 6)          # spectral
    ops.mlp_fused(x, w13, b13, w2, b2, resid, gamma=gamma, beta=beta)         # fused gated MLP (training: g kept)
    ops.mlp_fused(x, w13, b13, w2, b2, resid, gamma=gamma, beta=beta, keep_g=False)   # fused gated MLP (inference)
    ops.gemm_lnbwd(dab, w13t, xf, stats, gamma, dxf, dgamma=dgam, dbeta=dbet, inplace=True)    # dgrad K=1376 + LayerNorm backward
    ops.gemm_lnbwd(dqkv, wqkvt, xf, stats, gamma, dxf, dgamma=dgam, dbeta=dbet, inplace=True)  # dgrad K=768 + LayerNorm backward
    ops.wgrad_group(wg_jobs)                                                                   # the four weight gradients of a block
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
