"""CPU: host-side logic of the product, the C ABI surface, loud failure without CUDA."""
import ctypes
import hashlib
import os
import random
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, TINY
from oracle import hsimae_oracle as O


def test_swiglu_hidden_and_tables():
    from hsimae_b200.host import swiglu_hidden, sincos_table
    assert [swiglu_hidden(d) for d in (256, 128, 64, 144, 96, 32)] == [O.swiglu_hidden(d) for d in (256, 128, 64, 144, 96, 32)]
    assert (swiglu_hidden(256), swiglu_hidden(128), swiglu_hidden(64)) == (684, 344, 172)
    for w in (256, 128, 64, 32):
        assert torch.equal(sincos_table(w, 4, 3), O.sincos_table(w, 4, 3))


def test_visible_shape_matches_oracle_and_consumes_rng():
    from hsimae_b200.host import choose_visible_shape
    for ratio in (0.5, 0.75, 0.8, 0.9, 0.1, 0.25):
        for seed in range(6):
            random.seed(seed); a = choose_visible_shape(4, 9, ratio); ra = random.random()
            random.seed(seed); b = O.choose_visible_shape(4, 9, ratio); rb = random.random()
            assert a == b and ra == rb
    random.seed(0)
    assert {choose_visible_shape(4, 9, 0.5) for _ in range(40)} == {(2, 9), (3, 6)}
    assert {choose_visible_shape(4, 9, 0.8) for _ in range(40)} == {(2, 4), (4, 2)}
    assert {choose_visible_shape(4, 9, 0.75) for _ in range(10)} == {(3, 3)}
    with pytest.raises(ValueError):
        choose_visible_shape(1, 9, 0.5)


def test_init_hashes_match_reference_fixture(golden):
    import Models as M
    z = golden("init_hashes.npz")
    for cls, extra in (("HSIMAE", {}), ("DualViT", dict(num_class=17, drop_path=0.2)), ("HSIViT", dict(num_class=17))):
        kw = dict(TINY); kw.update(extra)
        if cls == "HSIViT":
            kw = {a: b for a, b in kw.items() if not a.startswith("decoder") and a != "norm_pix_loss"}
        torch.manual_seed(42); random.seed(42)
        m = getattr(M, cls)(**kw)
        assert "\n".join(m.state_dict().keys()) == str(z[cls + "_keys"])
        h = hashlib.sha1()
        for k, t in m.state_dict().items():
            h.update(k.encode()); h.update(t.numpy().tobytes())
        assert h.hexdigest() == str(z[cls])


def test_driver_contract_names():
    """the drivers split weight-decay groups on 'bias'/'norm' in parameter names (Model_Pretraining.py:80-84)"""
    import Models as M
    m = M.HSIMAE(**TINY)
    names = [n for n, _ in m.named_parameters()]
    assert "blocks_1.0.attn.q.weight" in names and "blocks_2.1.mlp.w3.bias" in names and "blocks.0.norm2.weight" in names
    assert "decoder_blocks.0.attn.proj.bias" in names and "decoder_pred.weight" in names and "mask_token" in names
    assert not m.pos_embed.requires_grad and not m.decoder_pos_embed.requires_grad
    assert m.blocks_1[0].mlp.w1.weight.shape == (172, 64)


def test_library_exports_every_declared_symbol():
    from hsimae_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hsimae_b200.h")).read()
    declared = set(re.findall(r"\b(hsimae_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    if not _lib.LIB_PATH.exists():
        from hsimae_b200.build import build
        build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    L = _lib.load()
    assert L.hsimae_abi_version() == 2


def test_plan_layout_is_consistent_with_module():
    """host-only plan queries (no device work): names, sizes, gradient arena buckets"""
    import Models as M
    from hsimae_b200 import _lib
    for cls, extra in ((M.HSIMAE, {}), (M.DualViT, dict(num_class=17)), (M.HSIViT, dict(num_class=17))):
        kw = dict(TINY); kw.update(extra)
        if cls is M.HSIViT:
            kw = {a: b for a, b in kw.items() if not a.startswith("decoder") and a != "norm_pix_loss"}
        m = cls(**kw)
        rt = m._runtime()
        named = dict(m.named_parameters())
        assert set(rt.names) == set(named) - {"mask_token"}
        assert all(named[n].numel() == k for n, k in zip(rt.names, rt.numels))
        with_grad = {n for n, o in zip(rt.names, rt.grad_off) if o >= 0}
        assert with_grad == {n for n, p in named.items() if p.requires_grad} - {"mask_token"}
        spans = sorted((o, o + k) for o, k in zip(rt.grad_off, rt.numels) if o >= 0)
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= rt.grad_elems
        assert sum(n for _, n in rt.buckets) == rt.grad_elems
        assert _lib.load().hsimae_plan_hidden(rt.plan, 0) == 172


def test_bad_dims_are_reported():
    from hsimae_b200 import _lib
    L = _lib.load()
    d = _lib.Dims(img_size=9, patch_size=3, bands=32, b_patch_size=8, embed_dim=100, depth=3, s_depth=2, num_heads=4,
                  dec_dim=32, dec_depth=1, dec_heads=4, num_class=0, qkv_bias=1, norm_pix_loss=1, mlp_ratio=4.0)
    h = ctypes.c_void_p()
    assert L.hsimae_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b"embed_dim" in L.hsimae_last_error()


def test_forward_without_cuda_fails_loudly():
    import Models as M
    m = M.HSIMAE(**TINY)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(2, 1, 32, 9, 9), mask_ratio=0.5)


def test_patch_layout_helpers_match_oracle():
    """HSIMAE.patchify / unpatchify / get_dim_patches (Models.py:461-493) are host-side layout / RNG helpers: any device"""
    import random
    import Models
    from conftest import TINY, tiny_geometry
    from oracle import hsimae_oracle as O
    m = Models.HSIMAE(**TINY)
    g = tiny_geometry()
    x = torch.randn(5, 1, 32, 9, 9)
    tok = m.patchify(x)
    assert tok.shape == (5, 36, 72) and torch.equal(tok, O.cube_to_patches(x, g))
    assert m.patch_info == (5, 32, 9, 9, 3, 8, 4, 3, 3)
    assert torch.equal(m.unpatchify(tok), x) and torch.equal(m.unpatchify(tok), O.patches_to_cube(tok, g))
    random.seed(3)
    lt, ll = m.get_dim_patches(4, 9, 0.5)
    random.seed(3)
    assert (int(lt), int(ll)) == O.choose_visible_shape(4, 9, 0.5) and lt.dtype == torch.int64 and lt.dim() == 0
    d = Models.DualViT(**{**TINY, "num_class": 5})
    assert torch.equal(d.patchify(x), tok)


def test_mean_var_attributes_follow_the_reference_definition():
    """`self.mean` / `self.var` (Models.py:605-610): per-patch mean and sqrt(unbiased var + 1e-6) of the last batch,
    evaluated lazily from the batch the last forward call saw (the fused loss kernel keeps them in registers)."""
    import Models
    m = Models.HSIMAE(**TINY)
    with pytest.raises(AttributeError):
        m.mean
    x = torch.randn(3, 1, 32, 9, 9)
    m.__dict__["_last_imgs"], m.__dict__["_last_stats"] = x, None     # what _masked_pass records
    t = m.patchify(x)
    assert torch.equal(m.mean, t.mean(dim=-1, keepdim=True)) and m.mean.shape == (3, 36, 1)
    assert torch.equal(m.var, (t.var(dim=-1, keepdim=True) + 1.0e-6) ** 0.5)


def test_second_backward_raises_instead_of_returning_zeros():
    """a second backward through one forward call must raise like torch does (the stash is released)."""
    from hsimae_b200 import modules as MM
    s = MM._new_saved()
    s.consumed = True

    class Ctx:
        saved = s
    with pytest.raises(RuntimeError, match="second time"):
        MM._HsiFunction.backward(Ctx(), None)
