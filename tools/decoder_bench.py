"""Device time of the decoder alone (hsimae_decoder_forward / hsimae_decoder_backward) at the bench shape: the baseline a
fused decoder has to beat.  Prints one JSON line.  `python tools/decoder_bench.py [batch]`
NOTE: written at the end of round 1 after the GPU budget was spent -- not yet run on hardware."""
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import Models  # noqa: E402
from bench import LARGE  # noqa: E402
from hsimae_b200 import _lib  # noqa: E402
from hsimae_b200.modules import _ptr, _stream  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0); random.seed(0)
model = Models.HSIMAE(**LARGE).cuda().train()
x = torch.randn(B, 1, 32, 9, 9, device="cuda")
rt, params = model._prepare(x)
T, Lp, lt, ll = 4, 9, 2, 9
noise_t, noise_l = torch.rand(B, T, device="cuda"), torch.rand(B, Lp, device="cuda")
ids_keep, ids_restore, mask, keep32, restore32 = rt.mask(noise_t, noise_l, lt, ll)
ws_enc = rt.enc_ws(B, lt, ll, True, x.device)
rt.encoder_forward(x, B, lt, ll, keep32, None, True, ws_enc)
ws_dec = rt.dec_ws(B, lt, ll, True, x.device)
loss = torch.empty(1, device="cuda"); pred = torch.empty_like(x); mimg = torch.empty_like(x)
grads = torch.zeros(rt.grad_elems, device="cuda")
g = torch.ones(1, device="cuda")


def fwd():
    _lib.check(rt.lib.hsimae_decoder_forward(rt.plan, _ptr(rt.wb), _ptr(rt.wf), _ptr(x), B, lt, ll, _ptr(restore32), _ptr(mask), _ptr(ws_enc),
                                             1, 1, _ptr(ws_dec), ws_dec.numel(), _ptr(loss), _ptr(pred), _ptr(mimg), None, _stream()), "decoder_forward")


def bwd():
    _lib.check(rt.lib.hsimae_decoder_backward(rt.plan, _ptr(rt.wb), _ptr(rt.wf), B, lt, ll, _ptr(restore32), _ptr(ws_enc), _ptr(ws_dec),
                                              ws_dec.numel(), _ptr(g), _ptr(grads), _stream()), "decoder_backward")


def timed(fn, reps=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        fwd() if fn is bwd else None          # backward consumes what the forward of the same step saved
        torch.cuda.synchronize()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


for _ in range(3):
    fwd(); bwd()
torch.cuda.synchronize()
l0 = rt.lib.hsimae_launch_count(); fwd(); l1 = rt.lib.hsimae_launch_count(); bwd(); l2 = rt.lib.hsimae_launch_count()
print(json.dumps({"batch": B, "decoder_rows": B * 36, "forward_ms": timed(fwd), "backward_ms": timed(bwd), "forward_launches": l1 - l0,
                  "backward_launches": l2 - l1, "loss": float(loss.item()),
                  "floor_ms_fused": {"forward": "8 x 37.7 MB block inputs written + in/out ~ 0.3", "backward": "~1.0 (DESIGN 4d)"}}))
