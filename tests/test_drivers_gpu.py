"""GPU box: the reference's UNCHANGED drivers -- `mask_pretraining` (Model_Pretraining.py:57-113),
`dual_branch_finetuning` (Model_Finetuning.py:66-240) and `test_model` (:243-301) -- run end to end against this
repository's drop-in `Models` module on cuda:0.  The driver sources come from the reference as mounted in the build
container, or (on the GPU box, where it is not mounted) from the byte-identical copy `oracle/fetch_ref.py` placed in
`oracle/_ref/`; the manifest check below asserts that copy is unmodified.  Shims are test-side only (see
test_drivers_cpu.py)."""
import os
from unittest import mock

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT
from test_drivers_cpu import _import_driver

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "Model_Pretraining.py")), reason="reference neither mounted nor fetched")]


def test_travelling_reference_copy_is_unmodified():
    from oracle import fetch_ref
    if REFERENCE.startswith(os.path.join(ROOT, "oracle")):
        assert fetch_ref.verify(), "oracle/_ref differs from its manifest"


def test_pretraining_driver_runs_unchanged_on_gpu(tmp_path):
    MP = _import_driver("Model_Pretraining")
    import Models
    assert MP.HSIMAE is Models.HSIMAE
    rng = np.random.default_rng(0)
    scene = rng.standard_normal((40, 41, 32))
    cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 32, 2) for w in range(0, 32, 2)], dtype=np.int16)   # 256 windows
    MP.seed_everything(42)
    MP.mask_pretraining([[scene], cut], str(tmp_path), "m.pkl", img_size=9, bands=32, mask_ratio=0.5, bs=64, epochs=6, depth=3,
                        dim=64, s_depth=2, dec_dim=32, dec_depth=1, lr=5e-3)
    sd = torch.load(tmp_path / "m.pkl")
    assert "blocks_1.0.attn.q.weight" in sd and "decoder_pred.bias" in sd and all(torch.isfinite(v).all() for v in sd.values())
    log = np.load(tmp_path / "train_log.npy", allow_pickle=True)
    losses = np.asarray(log[0], dtype=np.float64)
    assert losses.shape == (6,) and np.isfinite(losses).all()
    assert losses[-1] < losses[0], f"the reference loop did not reduce the loss: {losses}"


def test_finetuning_and_dense_inference_drivers_run_unchanged_on_gpu(tmp_path):
    MF = _import_driver("Model_Finetuning")
    import Models
    assert MF.DualViT is Models.DualViT and MF.HSIViT is Models.HSIViT
    rng = np.random.default_rng(1)
    H, W, C, n_class = 12, 11, 32, 4
    gt_map = rng.integers(0, n_class, size=(H, W))
    centres = rng.standard_normal((n_class, C)) * 2.0
    cubes = np.stack([centres[gt_map[i, j]][None, None, :] + 0.3 * rng.standard_normal((9, 9, C)) for i in range(H) for j in range(W)])
    labels = gt_map.reshape(-1)
    labeled_index = np.nonzero(labels)[0][:64]
    gt = labels[labeled_index]
    unlabeled = cubes[rng.permutation(len(cubes))[:60]]
    kw = dict(lr=1e-3, wd=5e-3, depth=3, dim=64, dec_depth=1, dec_dim=32, s_depth=2, epochs=4, mask_ratio=0.8, lamda=10, batch_size=8)
    it = torch.utils.data.dataloader._BaseDataLoaderIter
    with mock.patch.object(it, "next", it.__next__, create=True):
        MF.seed_everything(42)
        val_value, train_losses, val_losses = MF.dual_branch_finetuning(cubes, labeled_index, unlabeled, gt, str(tmp_path), "ft.pkl", **kw)
        assert len(train_losses) == 4 and np.isfinite(train_losses).all() and np.isfinite(val_losses).all()
        assert train_losses[-1] < train_losses[0]
        # dense per-pixel inference with the encoder-only model, loading the keys the fine-tuned checkpoint shares
        MF.mi = mock.MagicMock()
        oa, aa, kappa, ca = MF.test_model(cubes, gt_map, gt_map, str(tmp_path), "ft.pkl", depth=3, dim=64, s_depth=2)
    assert 0.0 <= oa <= 1.0 and np.isfinite(kappa)


def test_pretraining_loss_curve_matches_the_reference_run_on_cuda(tmp_path):
    """The SAME unchanged driver, seeds and data twice on cuda:0: once importing the reference's own Models.py (fp32
    eager, TF32 off), once this repository's drop-in.  Both consume torch's CUDA generator identically (mask noise), so
    the masks are equal and the per-epoch losses must agree to bf16-operand accuracy over 24 AdamW steps."""
    rng = np.random.default_rng(0)
    scene = rng.standard_normal((40, 41, 32))
    cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 32, 2) for w in range(0, 32, 2)], dtype=np.int16)
    kw = dict(img_size=9, bands=32, mask_ratio=0.5, bs=64, epochs=6, depth=3, dim=64, s_depth=2, dec_dim=32, dec_depth=1, lr=5e-3)
    curves = {}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for which in ("reference", "ours"):
            MP = _import_driver("Model_Pretraining", reference_models=(which == "reference"))
            out = tmp_path / which
            MP.seed_everything(42)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                MP.mask_pretraining([[scene], cut], str(out), "m.pkl", **kw)
            curves[which] = np.asarray(np.load(out / "train_log.npy", allow_pickle=True)[0], dtype=np.float64)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    import Models
    assert MP.HSIMAE is Models.HSIMAE
    rel = np.abs(curves["ours"] - curves["reference"]) / np.abs(curves["reference"])
    assert rel.max() < 5e-3, f"loss curves diverge: ours {curves['ours']} reference {curves['reference']}"
