#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q --tb=short -k attention 2>&1 | tail -3
timeout 600 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from hsimae_b200 import ops
B,D=4096,256
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/it*1e3
qkv=torch.randn(B*18,768,device='cuda').to(torch.bfloat16); do=torch.randn(B*18,256,device='cuda').to(torch.bfloat16)
for name,spec in (('fusion',(1,18,18,1)),('spatial',(3,6,6,1)),('spectral',(6,3,1,6))):
    out,lse=ops.attention_forward(qkv,B,D,16,18,*spec)
    print(name,'fwd us',round(t(lambda: ops.attention_forward(qkv,B,D,16,18,*spec)),1),'bwd us',round(t(lambda: ops.attention_backward(qkv,out,lse,do,B,D,16,18,*spec)),1))
qd=torch.randn(B*36,192,device='cuda').to(torch.bfloat16); dd=torch.randn(B*36,64,device='cuda').to(torch.bfloat16)
out,lse=ops.attention_forward(qd,B,64,8,36,1,36,36,1)
print('decoder fwd us',round(t(lambda: ops.attention_forward(qd,B,64,8,36,1,36,36,1)),1),'bwd us',round(t(lambda: ops.attention_backward(qd,out,lse,dd,B,64,8,36,1,36,36,1)),1))
PY
