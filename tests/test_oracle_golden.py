"""CPU: the oracle restatement reproduces what the reference computed (committed fixtures)."""
import hashlib

import numpy as np
import torch

from conftest import state_from_npz, grads_from_npz, tiny_geometry, rel_err
from oracle import hsimae_oracle as O


def test_pretrain_fixture(golden):
    z = golden("tiny_pretrain.npz")
    sd = state_from_npz(z)
    g = tiny_geometry()
    x = torch.from_numpy(z["x"])
    out, grads = O.pretrain_step_grads(sd, x, g, torch.from_numpy(z["noise_t"]), torch.from_numpy(z["noise_l"]),
                                       int(z["len_t"]), int(z["len_l"]))
    assert torch.equal(out["ids_keep"], torch.from_numpy(z["ids_keep"]))
    assert torch.equal(out["ids_restore"], torch.from_numpy(z["ids_restore"]))
    assert torch.equal(out["mask"], torch.from_numpy(z["mask_tokens"]))
    assert abs(out["loss"].item() - float(z["loss"])) < 2e-6
    assert torch.allclose(out["pred_img"], torch.from_numpy(z["pred"]), atol=2e-5)
    assert torch.equal(out["mask_img"], torch.from_numpy(z["mask"]))
    assert torch.allclose(out["latent"], torch.from_numpy(z["latent"]), atol=2e-5)
    ref = grads_from_npz(z)
    assert set(ref) == set(grads)
    for k in ref:
        assert rel_err(grads[k], ref[k]) < 1e-4, k


def test_dual_fixture(golden):
    z = golden("tiny_dual.npz")
    sd = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in state_from_npz(z).items()}
    g = tiny_geometry(17)
    xl, xu = torch.from_numpy(z["xl"]), torch.from_numpy(z["xu"])
    out = O.dual_forward(sd, xl, xu, g, torch.from_numpy(z["noise_t"]), torch.from_numpy(z["noise_l"]), int(z["len_t"]), int(z["len_l"]))
    total = 10.0 * out["loss"] + torch.nn.functional.cross_entropy(out["logits"], torch.from_numpy(z["labels"]), ignore_index=0)
    total.backward()
    assert abs(out["loss"].item() - float(z["loss_rec"])) < 2e-6
    assert abs(total.item() - float(z["total"])) < 2e-5
    assert torch.allclose(out["logits"], torch.from_numpy(z["logits"]), atol=2e-5)
    assert torch.allclose(out["pred_img"], torch.from_numpy(z["pred_rec"]), atol=2e-5)
    ref = grads_from_npz(z)
    for k, v in ref.items():
        assert rel_err(sd[k].grad, v) < 1e-4, k
    with torch.no_grad():
        ev = O.dual_forward(sd, xl, None, g)["logits"]
        assert torch.allclose(ev, torch.from_numpy(z["logits_eval"]), atol=2e-5)
        vit = O.vit_forward({k: v for k, v in sd.items() if not k.startswith("decoder") and k != "mask_token"}, xl, g)
        assert torch.allclose(vit, torch.from_numpy(z["logits_vit"]), atol=2e-5)


def test_kat_masks(golden):
    """index hashes of the seeded Base/Large runs."""
    z = golden("kat_masks.npz")
    nt, nl = torch.from_numpy(z["noise_t"]), torch.from_numpy(z["noise_l"])
    for name in ("base", "large"):
        lt, ll = (int(v) for v in z[name + "_shape"])
        assert (lt, ll) == (3, 6)
        ids_keep, ids_restore, mask = O.structured_mask(nt, nl, lt, ll)
        assert hashlib.sha1(ids_keep.numpy().tobytes()).hexdigest()[:16] == "5ee5fac47baaef6e"
        assert hashlib.sha1(ids_restore.numpy().tobytes()).hexdigest()[:16] == "466597abc0b24641"
        assert np.array_equal(mask.numpy(), z[name + "_mask"])
    assert ids_keep[0].tolist() == [1, 2, 4, 5, 6, 7, 10, 11, 13, 14, 15, 16, 28, 29, 31, 32, 33, 34]


def test_mask_ties_and_shapes():
    """forced ties resolve lowest-index-first; every legal visible shape is a valid permutation"""
    torch.manual_seed(0)
    for T, L in ((4, 9), (2, 4), (8, 16)):
        for lt in range(1, T + 1):
            for ll in range(1, L + 1):
                nt, nl = torch.rand(5, T), torch.rand(5, L)
                nt[:, 1] = nt[:, 0]
                nl[:, -1] = nl[:, 0]
                k, r, m = O.structured_mask(nt, nl, lt, ll)
                assert torch.equal(torch.sort(r, 1).values, torch.arange(T * L).expand(5, -1))
                assert int(m.sum()) == 5 * (T * L - lt * ll)
                assert torch.equal(torch.sort(k, 1).values, k)
                assert bool((m.gather(1, k) == 0).all())
