#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "attention" 2>&1 | tail -4
HSIMAE_ATTN_SMALL=0 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "attention" 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch, test_ops_gpu as T
from conftest import rel_err
from hsimae_b200 import ops
for n, D, heads, lt, ll, kind in [(37, 256, 16, 3, 6, "spatial"), (37, 256, 16, 3, 6, "full"), (20, 256, 16, 2, 9, "spatial"), (300, 64, 8, 4, 9, "full"), (9, 128, 8, 4, 9, "spatial")]:
    K = lt * ll
    if kind == "spatial": spec = (lt, ll, ll, 1); groups = [torch.arange(ll) + t * ll for t in range(lt)]
    else: spec = (1, K, K, 1); groups = [torch.arange(K)]
    groups = [g.to("cuda") for g in groups]
    qkv = T._rand_bf16(n * K, 3 * D, seed=18)
    out, lse = ops.attention_forward(qkv, n, D, heads, K, *spec)
    leaf = qkv.float().requires_grad_(True)
    ref = T._attn_ref(leaf, n, D, heads, K, groups)
    dout = T._rand_bf16(n * K, D, scale=0.1, seed=19)
    ref.backward(dout.float())
    dqkv = ops.attention_backward(qkv, out, lse, dout, n, D, heads, K, *spec).float()
    g = leaf.grad
    print(kind, K, D, "rel err dq %.2e dk %.2e dv %.2e" % (rel_err(dqkv[:, :D], g[:, :D]), rel_err(dqkv[:, D:2*D], g[:, D:2*D]), rel_err(dqkv[:, 2*D:], g[:, 2*D:])))
PY
python tools/attn_bench.py 2>&1 | tail -1
