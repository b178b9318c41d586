import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsimae_b200 import ops
B, D = 4096, 256
qkv = torch.randn(B * 18, 768, device='cuda').to(torch.bfloat16); do = torch.randn(B * 18, 256, device='cuda').to(torch.bfloat16)
for _ in range(2):
    out, lse = ops.attention_forward(qkv, B, D, 16, 18, 1, 18, 18, 1)
    ops.attention_backward(qkv, out, lse, do, B, D, 16, 18, 1, 18, 18, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out, lse = ops.attention_forward(qkv, B, D, 16, 18, 1, 18, 18, 1)
ops.attention_backward(qkv, out, lse, do, B, D, 16, 18, 1, 18, 18, 1)
out, lse = ops.attention_forward(qkv, B, D, 16, 18, 6, 3, 1, 6)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
