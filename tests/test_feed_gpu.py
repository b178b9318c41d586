"""GPU: the on-device data feed (hsimae_gather_patches through hsimae_b200.feed.PatchFeed) against the oracle and the
reference-generated fixture.  Pure data movement + one IEEE subtract / divide: the bar is bit-exact."""
import random

import numpy as np
import pytest
import torch

from oracle import feed_oracle as FO

pytestmark = pytest.mark.gpu


def _feed(scenes, cut, train, img=9):
    from hsimae_b200.feed import PatchFeed
    return PatchFeed([scenes, cut], train=train, device="cuda:0", img=img)


def test_items_match_reference_fixture(golden):
    z = golden("feed.npz")
    scenes, cut = [z["scene0"], z["scene1"]], z["cut_info"]
    feed = _feed(scenes, cut, train=True)
    random.seed(123)
    got = feed.batch(np.arange(len(cut)))                          # draws the flips like the reference
    assert got.shape == (len(cut), 1, 32, 9, 9) and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), z["train_items"])
    assert np.array_equal(_feed(scenes, cut, train=False).batch(np.arange(len(cut))).cpu().numpy(), z["eval_items"])


def test_seeded_epoch_matches_reference_dataloader(golden):
    z = golden("feed.npz")
    feed = _feed([z["scene0"], z["scene1"]], z["cut_info"], train=True)
    torch.manual_seed(7); random.seed(7)
    got = torch.cat(list(feed.epoch(batch_size=5, shuffle=True)))
    assert np.array_equal(got.cpu().numpy(), z["epoch_batches"])


@pytest.mark.parametrize("n_scenes,bands,img,batch", [(3, 32, 9, 4096), (1, 32, 9, 1), (2, 16, 5, 257), (2, 8, 11, 33)])
def test_random_batches_bit_exact(n_scenes, bands, img, batch):
    rng = np.random.default_rng(n_scenes * 100 + bands)
    scenes = [(rng.standard_normal((img + int(rng.integers(0, 40)), img + int(rng.integers(0, 50)), bands)) * (k + 1)).astype(np.float32)
              for k in range(n_scenes)]
    rows = []
    for k, s in enumerate(scenes):
        mx, mn = (1, 0) if k % 2 == 0 else (int(s.max()) + 1, int(s.min()) - 1)
        for _ in range(60):
            rows.append((0, int(rng.integers(0, s.shape[0] - img + 1)), int(rng.integers(0, s.shape[1] - img + 1)), k, mx, mn))
        rows.append((0, s.shape[0] - img, s.shape[1] - img, k, mx, mn))     # window touching the far corner
        rows.append((0, 0, 0, k, mx, mn))
    cut = np.array(rows, dtype=np.int16)
    idx = rng.integers(0, len(cut), size=batch)
    flips = rng.integers(0, 2, size=(batch, 2)).astype(np.uint8)
    feed = _feed(scenes, cut, train=True, img=img)
    got = feed.batch(idx, torch.from_numpy(flips)).cpu().numpy()
    ref = FO.get_batch(scenes, cut, idx, flips, img=img)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    # full-size property: flipping twice is the identity, flips only permute values inside each band plane
    noflip = feed.batch(idx, torch.zeros(batch, 2, dtype=torch.uint8)).cpu().numpy()
    assert np.array_equal(np.sort(got.reshape(batch, bands, -1), axis=2), np.sort(noflip.reshape(batch, bands, -1), axis=2))


def test_edges_and_errors():
    from hsimae_b200.feed import PatchFeed
    s = np.zeros((9, 9, 32), dtype=np.float32)
    cut = np.array([(0, 0, 0, 0, 1, 0)], dtype=np.int16)
    feed = PatchFeed([[s], cut], train=True)
    assert feed.batch([]).shape == (0, 1, 32, 9, 9)
    with pytest.raises(IndexError):
        feed.batch([1])
    with pytest.raises(ValueError):
        PatchFeed([[s], np.array([(0, 1, 0, 0, 1, 0)], dtype=np.int16)])      # window leaves the scene
    with pytest.raises(ValueError):
        PatchFeed([[s], np.array([(0, 0, 0, 0, 3, 3)], dtype=np.int16)])      # max == min
    with pytest.raises(RuntimeError):
        PatchFeed([[s], cut], device="cpu")


def test_feeds_the_model():
    """A pretraining step straight from the device feed (what Model_Pretraining.py:96-101 does with DataLoader batches)."""
    import Models
    rng = np.random.default_rng(0)
    scene = rng.standard_normal((40, 40, 32)).astype(np.float32)
    cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 32, 3) for w in range(0, 32, 3)], dtype=np.int16)
    feed = _feed([scene], cut, train=True)
    torch.manual_seed(0); random.seed(0)
    model = Models.HSIMAE(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=128, depth=12, num_heads=8, s_depth=9,
                          decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True).to("cuda:0")
    losses = []
    for x in feed.epoch(batch_size=32):
        loss, pred, mask = model(x, mask_ratio=0.5)
        loss.backward()
        losses.append(loss.item())
    assert len(losses) == (len(cut) + 31) // 32 and all(np.isfinite(losses))


def test_raw_scene_to_batches_pipeline(tmp_path):
    """raw .npy scenes -> device group-wise PCA -> cut table -> device batches (Utils/Preprocessing.py:82-118 +
    Model_Pretraining.py:40-51,76), against the oracles of each stage."""
    from hsimae_b200.feed import PatchFeed, get_data_cut_file, split_info
    from hsimae_b200.gwpca import _auto_sign
    from oracle import feed_oracle as FO, gwpca_oracle as G
    rng = np.random.default_rng(3)
    paths, raw = [], []
    for i, (h, w) in enumerate(((20, 26), (15, 13))):
        x = rng.normal(size=(h, w, 6)) @ (rng.normal(size=(6, 120)) * np.linspace(3, 0.5, 6)[:, None])
        x = np.round((x + 0.2 * rng.normal(size=(h, w, 120))) * 400.0 + 6000.0)
        np.save(tmp_path / f"s{i}.npy", x)
        paths.append(str(tmp_path / f"s{i}.npy")); raw.append(x)
    np.random.seed(11)
    cubes, cut = get_data_cut_file(paths, patch_size=9, norm=False, GWPCA=True, ratio=0.75)
    # stage oracles: PCA per scene, then the cut table with the same numpy-RNG shuffles
    ref_cubes = [G.apply_gwpca(x, 32, 4, True, _auto_sign()) for x in raw]
    np.random.seed(11)
    ref_cut = []
    for k, rc in enumerate(ref_cubes):
        rows = np.array(split_info(rc.shape, (9, 9, 32), (3, 3, 1), k, 1, 0))
        np.random.shuffle(rows)
        ref_cut += list(rows[:int(rows.shape[0] * 0.75)])
    ref_cut = np.array(ref_cut, dtype=np.int16)
    assert cut.dtype == np.int16 and np.array_equal(cut, ref_cut)
    for c, rc in zip(cubes, ref_cubes):
        assert c.is_cuda and c.dtype == torch.float64 and np.abs(c.cpu().numpy() - rc).max() <= 1e-6 * np.abs(rc).max()
    feed = PatchFeed([cubes, cut], train=True)
    idx = list(range(len(cut)))[::-1]
    random.seed(5)
    got = feed.batch(idx)
    random.seed(5)
    want = FO.get_batch(ref_cubes, ref_cut, idx, FO.draw_flips(len(idx)))
    assert got.shape == want.shape and np.abs(got.cpu().numpy() - want).max() <= 1e-5     # fp32 cubes of O(1) whitened values
