"""CPU, build container only: pin the oracle and the host-side boundary against the LIVE reference."""
import contextlib
import importlib.util
import io
import os
import random

import pytest
import torch

from conftest import REFERENCE, TINY, tiny_geometry, rel_err
from oracle import hsimae_oracle as O

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "Models.py")), reason="reference not mounted")


@pytest.fixture(scope="module")
def R():
    spec = importlib.util.spec_from_file_location("_reference_models", os.path.join(REFERENCE, "Models.py"))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("ratio", [0.5, 0.75, 0.8, 0.9, 0.25])
def test_pretrain_matches_reference(R, ratio):
    torch.manual_seed(3); random.seed(3)
    m = quiet(R.HSIMAE, **TINY)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = tiny_geometry()
    torch.manual_seed(11); random.seed(11)
    x = torch.randn(5, 1, 32, 9, 9)
    st = (torch.get_rng_state(), random.getstate())
    loss, pred, mask = m(x, mask_ratio=ratio)
    loss.backward()
    torch.set_rng_state(st[0]); random.setstate(st[1])
    lt, ll = O.choose_visible_shape(4, 9, ratio)
    assert (lt, ll) == (int(m.len_t), int(m.len_l))
    out, grads = O.pretrain_step_grads(sd, x, g, torch.rand(5, 4), torch.rand(5, 9), lt, ll)
    assert abs(out["loss"].item() - loss.item()) < 2e-6
    assert torch.allclose(out["pred_img"], pred, atol=2e-5) and torch.equal(out["mask_img"], mask)
    for k, p in m.named_parameters():
        if p.grad is None:
            assert k not in grads
        else:
            assert rel_err(grads[k], p.grad) < 1e-4, k


def test_host_rng_contract_matches_reference(R):
    """product host code (visible-shape choice, drop draws) consumes RNG exactly like the reference"""
    from hsimae_b200.host import choose_visible_shape, draw_drop_factors
    kw = dict(TINY); kw.update(num_class=17, drop_path=0.3)
    torch.manual_seed(8); random.seed(8)
    d = quiet(R.DualViT, **kw)
    d.train()
    sd = {k: v.detach().clone() for k, v in d.state_dict().items()}
    g = tiny_geometry(17)
    rates = [b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0 for b in d.blocks_1]
    rates_f = [b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0 for b in d.blocks]
    torch.manual_seed(21); random.seed(21)
    xl, xu = torch.randn(4, 1, 32, 9, 9), torch.randn(6, 1, 32, 9, 9)
    st = (torch.get_rng_state(), random.getstate())
    loss, pred, mask, logits = d(xl, xu, mask_ratio=0.8)
    torch.set_rng_state(st[0]); random.setstate(st[1])
    drops_full = draw_drop_factors(rates, rates_f, 4, 4, 9, "cpu", True)
    lt, ll = choose_visible_shape(4, 9, 0.8)
    nt, nl = torch.rand(10, 4), torch.rand(10, 9)
    drops_m = draw_drop_factors(rates, rates_f, 10, lt, ll, "cpu", True)

    def as_dict(lst, sdepth=2):
        out = {}
        for stack, base in ((1, 0), (2, 2 * sdepth), (0, 4 * sdepth)):
            n = sdepth if stack else (len(lst) - 4 * sdepth) // 2
            for i in range(n):
                out[(stack, i, 1)] = lst[base + 2 * i]
                out[(stack, i, 2)] = lst[base + 2 * i + 1]
        return out

    out = O.dual_forward(sd, xl, xu, g, nt, nl, lt, ll, as_dict(drops_full), as_dict(drops_m))
    assert abs(out["loss"].item() - loss.item()) < 2e-6
    assert torch.allclose(out["logits"], logits, atol=2e-5)
    assert torch.equal(out["mask_img"], mask)


@pytest.mark.parametrize("cls,extra", [("HSIMAE", {}), ("DualViT", dict(num_class=17, drop_path=0.2)), ("HSIViT", dict(num_class=17))])
def test_module_init_matches_reference(R, cls, extra):
    import Models as M
    kw = dict(TINY); kw.update(extra)
    if cls == "HSIViT":
        kw = {a: b for a, b in kw.items() if not a.startswith("decoder") and a != "norm_pix_loss"}
    for trunc in (True, False):
        kw["trunc_init"] = trunc
        torch.manual_seed(42); random.seed(42)
        r = quiet(getattr(R, cls), **kw)
        a = torch.rand(2)
        torch.manual_seed(42); random.seed(42)
        m = getattr(M, cls)(**kw)
        b = torch.rand(2)
        sr, sm = r.state_dict(), m.state_dict()
        assert list(sr) == list(sm)
        assert all(torch.equal(sr[k], sm[k]) for k in sr)
        assert torch.equal(a, b)
        assert [n for n, p in r.named_parameters() if not p.requires_grad] == [n for n, p in m.named_parameters() if not p.requires_grad]
        m.load_state_dict(r.state_dict())


def test_patch_layout_helpers_match_reference(R):
    """patchify / unpatchify / get_dim_patches of the drop-in classes against the reference's own methods"""
    import Models
    torch.manual_seed(0)
    ref = quiet(R.HSIMAE, **TINY)
    ours = Models.HSIMAE(**TINY)
    x = torch.randn(6, 1, 32, 9, 9)
    a, b = ref.patchify(x), ours.patchify(x)
    assert torch.equal(a, b) and tuple(ref.patch_info) == tuple(ours.patch_info)
    assert torch.equal(ref.unpatchify(a), ours.unpatchify(b)) and torch.equal(ours.unpatchify(b), x)
    for ratio in (0.25, 0.5, 0.75, 0.8, 0.9):
        for seed in range(4):
            random.seed(seed); ra = ref.get_dim_patches(4, 9, ratio); sa = random.getstate()
            random.seed(seed); rb = ours.get_dim_patches(4, 9, ratio); sb = random.getstate()
            assert (int(ra[0]), int(ra[1])) == (int(rb[0]), int(rb[1])) and sa == sb
            assert rb[0].dtype == ra[0].dtype and rb[0].dim() == ra[0].dim()
