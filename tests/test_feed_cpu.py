"""CPU: the data-feed oracle against fixtures produced by the reference's own HSIdataset4PT / DataLoader
(tests/golden/feed.npz, oracle/make_golden_feed.py), and the host-side order / flip logic of hsimae_b200.feed."""
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE
from oracle import feed_oracle as FO


def _data(z):
    return [z["scene0"], z["scene1"]], z["cut_info"]


def test_oracle_items_match_reference_fixture(golden):
    z = golden("feed.npz")
    scenes, cut = _data(z)
    random.seed(123)
    flips = FO.draw_flips(len(cut), train=True)
    got = np.stack([FO.get_item(scenes, cut, i, flips[i]) for i in range(len(cut))])
    assert got.dtype == np.float32 and got.shape == z["train_items"].shape
    assert np.array_equal(got, z["train_items"])                      # bit-exact, flips included
    assert flips.any() and not flips.all()
    assert np.array_equal(FO.get_batch(scenes, cut, np.arange(len(cut))), z["eval_items"])


def test_epoch_order_and_flips_match_reference_dataloader(golden):
    from hsimae_b200.feed import loader_order
    z = golden("feed.npz")
    scenes, cut = _data(z)
    torch.manual_seed(7); random.seed(7)
    order = loader_order(len(cut), shuffle=True)
    assert sorted(order.tolist()) == list(range(len(cut)))
    got = []
    for i in range(0, len(cut), 5):
        idx = order[i:i + 5].numpy()
        got.append(FO.get_batch(scenes, cut, idx, FO.draw_flips(len(idx))))
    assert np.array_equal(np.concatenate(got), z["epoch_batches"])


def test_loader_order_without_shuffle_consumes_base_seed_only():
    from hsimae_b200.feed import loader_order
    torch.manual_seed(3)
    assert loader_order(6, shuffle=False).tolist() == list(range(6))
    after = torch.rand(1)
    torch.manual_seed(3)
    torch.empty((), dtype=torch.int64).random_()
    assert torch.equal(after, torch.rand(1))


def test_oracle_edges(golden):
    z = golden("feed.npz")
    scenes, cut = _data(z)
    assert FO.get_batch(scenes, cut, []).shape == (0, 1, 32, 9, 9)
    assert FO.draw_flips(4, train=False).sum() == 0
    # normalisation uses the int16-truncated max / min of the cut table (Utils/Preprocessing.py:114)
    i = int(np.nonzero(cut[:, 3] == 1)[0][0])
    c, h, w, num, mx, mn = cut[i]
    ref = (scenes[1][h:h + 9, w:w + 9, :] - mn) / (mx - mn)
    assert np.array_equal(FO.get_item(scenes, cut, i)[0], np.transpose(ref.astype(np.float32), (2, 0, 1)))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference sources not mounted")
def test_oracle_matches_live_reference_dataset():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    try:
        import make_golden_feed as G
    finally:
        sys.path.pop(0)
    MP = G.import_reference_pretraining()
    scenes, cut = G.synthetic(seed=5)
    ds = MP.HSIdataset4PT([scenes, cut], train=True)
    random.seed(99)
    ref = np.stack([ds[i].numpy() for i in range(len(ds))])
    random.seed(99)
    flips = FO.draw_flips(len(cut))
    assert np.array_equal(FO.get_batch(scenes, cut, np.arange(len(cut)), flips), ref)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_cut_table_matches_live_reference(tmp_path):
    """`get_data_cut_file` (Utils/Preprocessing.py:82-118) without the PCA step: same int16 cut table, incl. the numpy-RNG
    row shuffle, the `ratio` cut, the stride switch at scene 14 and the int16 truncation of max / min."""
    sys.path.insert(0, REFERENCE)
    try:
        from Utils.Preprocessing import get_data_cut_file as ref_cut, get_inital_seq
    finally:
        sys.path.remove(REFERENCE)
    from hsimae_b200.feed import get_data_cut_file, initial_seq
    for length, size, stride in ((9, 9, 3), (10, 9, 3), (11, 9, 3), (12, 9, 3), (31, 9, 3), (27, 9, 1), (28, 9, 1), (32, 32, 1), (217, 9, 3)):
        assert np.array_equal(initial_seq(length, size, stride), get_inital_seq(length, size, stride))
    rng = np.random.default_rng(0)
    paths = []
    for i, (h, w) in enumerate([(14, 19), (9, 9), (23, 12)] + [(10, 11)] * 12 + [(21, 30)]):    # 16 scenes: the last two take the stride-1 branch
        p = tmp_path / f"s{i}.npy"
        np.save(p, rng.normal(size=(h, w, 32)) * 3.0 + 1.5)
        paths.append(str(p))
    for norm, ratio in ((False, 1.0), (True, 0.6)):
        np.random.seed(7)
        ref = ref_cut(paths, patch_size=9, norm=norm, GWPCA=False, ratio=ratio)
        np.random.seed(7)
        ours = get_data_cut_file(paths, patch_size=9, norm=norm, GWPCA=False, ratio=ratio)
        assert ours[1].dtype == np.int16 and np.array_equal(ours[1], ref[1])
        assert all(np.array_equal(a, b) for a, b in zip(ours[0], ref[0]))
