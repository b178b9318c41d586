"""How much of the pretraining step is kernel-boundary cost?  Forward + backward of HSIMAE-Large (batch 4096) replayed from ONE
CUDA graph vs the eager launches (tuning probe; the headline metric keeps the driver's eager loop)."""
import os, sys, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import Models
import hsimae_b200.modules as mod

LARGE = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=256, depth=12, num_heads=16, s_depth=9,
             decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
torch.manual_seed(42); random.seed(42)
model = Models.HSIMAE(**LARGE).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-3, weight_decay=5e-2, betas=(0.9, 0.95))
B = 4096
x = torch.randn(B, 1, 32, 9, 9, device="cuda")

def fwd_bwd():
    loss, _, _ = model(x, mask_ratio=0.5)
    loss.backward()
    return loss

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def eager():
    opt.zero_grad(); fwd_bwd(); opt.step()
t_eager = timeit(eager)
def eager_nostep():
    opt.zero_grad(); fwd_bwd()
t_eager_fb = timeit(eager_nostep)

mod.choose_visible_shape = lambda T, L, r: (3, 6)
for _ in range(2):
    opt.zero_grad(set_to_none=True); fwd_bwd()
torch.cuda.synchronize()
opt.zero_grad(set_to_none=True)
model._prepare(x); model.invalidate_weight_cache()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = fwd_bwd()
def graphed():
    g.replay(); opt.step()
t_graph = timeit(graphed)
t_graph_fb = timeit(lambda: g.replay())
print(json.dumps({"eager_ms": t_eager, "eager_fwd_bwd_ms": t_eager_fb, "graph_ms": t_graph, "graph_fwd_bwd_ms": t_graph_fb, "loss": float(loss)}))
