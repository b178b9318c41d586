"""GPU box: this repository's drop-in classes against the UNMODIFIED reference classes, both on cuda:0.

The reference's Models.py (mounted, or the byte-identical travelling copy of oracle/fetch_ref.py) runs in fp32 eager
mode with TF32 off; the product runs its sm_100a path.  Same seeds => both consume Python's `random` and torch's CUDA
generator identically (Models.py:490,506,513), so the masks must be EQUAL bit for bit from the seed alone -- no noise
is injected -- and loss / reconstruction / logits / gradients must agree within the tolerances of test_model_gpu.py."""
import os
import random

import pytest
import torch
import torch.nn.functional as F

from conftest import REFERENCE, TINY, rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "Models.py")), reason="reference neither mounted nor fetched")]
DEV = "cuda"


@pytest.fixture(scope="module")
def R():
    from oracle import fetch_ref
    return fetch_ref.import_models()


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _pair(R, cls, kw, seed):
    import Models as M
    import contextlib, io
    torch.manual_seed(seed); random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = getattr(R, cls)(**kw)
    torch.manual_seed(seed); random.seed(seed)
    ours = getattr(M, cls)(**kw)
    for (ka, a), (kb, b) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert ka == kb and torch.equal(a, b), ka          # seeded construction is bit-identical
    return ref.to(DEV), ours.to(DEV)


def _grad_check(ref, ours, tol=6e-2, tol_vec=1e-1):
    named = dict(ours.named_parameters())
    bad = []
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert named[k].grad is None, k
            continue
        if k.endswith("attn.k.bias"):   # exactly zero in real arithmetic (softmax shift invariance): rounding noise on both sides
            continue
        e = rel_err(named[k].grad, p.grad)
        if not e < (tol_vec if p.dim() == 1 else tol):
            bad.append((k, e))
    assert not bad, "gradient mismatch vs the reference on CUDA: " + ", ".join(f"{k}:{e:.3g}" for k, e in bad[:10])


@pytest.mark.parametrize("ratio", [0.5, 0.75, 0.8])
def test_hsimae_seed_only_masks_and_outputs_match_reference_on_cuda(R, ratio):
    ref, ours = _pair(R, "HSIMAE", TINY, seed=5)
    x = torch.randn(24, 1, 32, 9, 9, device=DEV)
    torch.manual_seed(21); random.seed(21)
    l0, p0, m0 = ref(x, mask_ratio=ratio)
    l0.backward()
    torch.manual_seed(21); random.seed(21)
    l1, p1, m1 = ours(x, mask_ratio=ratio)
    l1.backward()
    assert torch.equal(m0, m1), "mask differs from the reference under the same seed"
    assert (int(ref.len_t), int(ref.len_l)) == (int(ours.len_t), int(ours.len_l))
    assert abs(l1.item() - l0.item()) <= 3e-3 * abs(l0.item())
    assert rel_err(p1, p0) < 2e-2
    assert rel_err(ours.mean, ref.mean) < 1e-6 and rel_err(ours.var, ref.var) < 1e-5
    _grad_check(ref, ours)


@pytest.mark.parametrize("drop_path", [0.0, 0.2])
def test_dualvit_training_step_matches_reference_on_cuda(R, drop_path):
    """with drop_path > 0 the stochastic-depth draws (Models.py:244-251) precede the mask noise on the CUDA generator:
    equal masks prove the draw order / shapes match the reference, equal outputs that the factors are applied alike"""
    kw = {**TINY, "num_class": 6, "drop_path": drop_path}
    ref, ours = _pair(R, "DualViT", kw, seed=9)
    ref.train(); ours.train()
    xl, xu = torch.randn(8, 1, 32, 9, 9, device=DEV), torch.randn(13, 1, 32, 9, 9, device=DEV)
    y = torch.randint(0, 6, (8,), device=DEV)
    outs = []
    for m in (ref, ours):
        torch.manual_seed(33); random.seed(33)
        loss_rec, pred, mask, logits = m(xl, xu, mask_ratio=0.8)
        (10.0 * loss_rec + F.cross_entropy(logits, y, ignore_index=0)).backward()
        outs.append((loss_rec, pred, mask, logits))
    (l0, p0, m0, g0), (l1, p1, m1, g1) = outs
    assert torch.equal(m0, m1)
    assert abs(l1.item() - l0.item()) <= 3e-3 * abs(l0.item())
    assert rel_err(p1, p0) < 2e-2 and rel_err(g1, g0) < 2e-2
    _grad_check(ref, ours)
    ref.eval(); ours.eval()
    with torch.no_grad():
        assert rel_err(ours(xl), ref(xl)) < 2e-2


def test_hsivit_logits_match_reference_on_cuda(R):
    kw = {k: v for k, v in TINY.items() if not k.startswith("decoder") and k != "norm_pix_loss"}
    kw["num_class"] = 6
    ref, ours = _pair(R, "HSIViT", kw, seed=3)
    ref.eval(); ours.eval()
    x = torch.randn(50, 1, 32, 9, 9, device=DEV)
    with torch.no_grad():
        assert rel_err(ours(x), ref(x)) < 2e-2
