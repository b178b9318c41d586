"""CPU, build container only: the reference's UNCHANGED training drivers (Model_Pretraining.mask_pretraining,
Model_Finetuning.dual_branch_finetuning) run against this repository's drop-in `Models` module up to the first model
call -- model construction with the drivers' keywords, `.to(device)`, the 'bias'/'norm' weight-decay split, AdamW, the
cosine schedule, datasets / loaders, the first batch -- where, without a GPU, the product path must refuse loudly
(there is no CPU fallback).  With a GPU *and* the reference present the loops run to completion instead.

Test-side shims only: `timm` (absent: our restated CosineLRScheduler), `matplotlib` (absent: mocks),
`_BaseDataLoaderIter.next` (removed from torch after 1.12), ragged `np.array([[...], []])` (an error since numpy 1.24;
the drivers were written for 1.2x where it made an object array -- Model_Pretraining.py:112).  The driver sources are
imported as they are."""
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "Model_Pretraining.py")), reason="reference not mounted")


class _NumpyCompat:
    """the driver's `np`: numpy, except that a ragged `np.array(...)` builds an object array as numpy < 1.24 did"""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, k):
        return getattr(self._real, k)

    def array(self, obj, *a, **k):
        try:
            return self._real.array(obj, *a, **k)
        except ValueError:
            return self._real.array(obj, *a, dtype=object, **k)


def _import_driver(name, reference_models=False):
    """import the reference's driver module `name` unchanged; its `from Models import ...` resolves to this repository's
    drop-in (default) or, with reference_models=True, to the reference's own Models.py (the fp32 baseline run)"""
    from hsimae_b200.optim import CosineLRScheduler
    timm, sched = types.ModuleType("timm"), types.ModuleType("timm.scheduler")
    sched.CosineLRScheduler = CosineLRScheduler
    timm.scheduler = sched
    shims = {"timm": timm, "timm.scheduler": sched, "matplotlib": mock.MagicMock(), "matplotlib.pyplot": mock.MagicMock(),
             "matplotlib.image": mock.MagicMock()}
    import Models                                   # this repository's drop-in (ROOT is first on sys.path, see conftest)
    assert os.path.dirname(os.path.abspath(Models.__file__)) == ROOT
    sys.modules.pop(name, None)
    if reference_models:
        from oracle import fetch_ref
        shims["Models"] = fetch_ref.import_models()
    with mock.patch.dict(sys.modules, shims):
        sys.path.insert(1, REFERENCE)               # behind ROOT: `from Models import ...` resolves to ours
        try:
            mod = __import__(name)
        finally:
            sys.path.remove(REFERENCE)
    sys.modules.pop(name, None)
    mod.CosineLRScheduler = CosineLRScheduler
    mod.tqdm = lambda it, *a, **k: it
    mod.np = _NumpyCompat(np)
    return mod


def test_pretraining_driver_reaches_the_model(tmp_path):
    MP = _import_driver("Model_Pretraining")
    import Models
    assert MP.HSIMAE is Models.HSIMAE
    rng = np.random.default_rng(0)
    scene = rng.standard_normal((15, 18, 32))
    cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 7, 3) for w in range(0, 10, 3)], dtype=np.int16)
    kw = dict(img_size=9, bands=32, mask_ratio=0.5, bs=4, epochs=1, depth=3, dim=64, s_depth=2, dec_dim=32, dec_depth=1)
    if torch.cuda.is_available():
        MP.mask_pretraining([[scene], cut], str(tmp_path), "m.pkl", **kw)
        sd = torch.load(tmp_path / "m.pkl")
        assert "blocks_1.0.attn.q.weight" in sd and "decoder_pred.bias" in sd
    else:
        with pytest.raises(RuntimeError, match="hsimae_b200 runs on a CUDA"):
            MP.mask_pretraining([[scene], cut], str(tmp_path), "m.pkl", **kw)


def test_finetuning_driver_reaches_the_model(tmp_path):
    MF = _import_driver("Model_Finetuning")
    import Models
    assert MF.DualViT is Models.DualViT and MF.HSIViT is Models.HSIViT
    rng = np.random.default_rng(1)
    n_lab, n_unl, n_class = 24, 30, 4
    data_list = rng.standard_normal((n_lab + 10, 9, 9, 32))
    labeled_index = list(range(n_lab))
    gt = np.array([1 + i % (n_class - 1) for i in range(n_lab)])
    unlabeled = rng.standard_normal((n_unl, 9, 9, 32))
    kw = dict(lr=1e-3, wd=5e-3, depth=3, dim=64, dec_depth=1, dec_dim=32, s_depth=2, epochs=1, mask_ratio=0.8, lamda=10, batch_size=4)
    it = torch.utils.data.dataloader._BaseDataLoaderIter
    with mock.patch.object(it, "next", it.__next__, create=True):
        if torch.cuda.is_available():
            MF.dual_branch_finetuning(data_list, labeled_index, unlabeled, gt, str(tmp_path), "ft.pkl", **kw)
        else:
            with pytest.raises(RuntimeError, match="hsimae_b200 runs on a CUDA"):
                MF.dual_branch_finetuning(data_list, labeled_index, unlabeled, gt, str(tmp_path), "ft.pkl", **kw)
