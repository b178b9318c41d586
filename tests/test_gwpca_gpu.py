"""GPU: group-wise PCA preprocessing (hsimae_b200.gwpca.applyGWPCA -> hsimae_gwpca_moments / hsimae_gwpca_project)
against the CPU oracle and the fixture produced by the reference's applyGWPCA (Utils/GroupWisePCA.py:20-34).

Tolerance: all device arithmetic is fp64; the oracle / sklearn differ only in summation order and LAPACK driver, so
max |d| <= 1e-6 * max |ref| (measured ~1e-9).  Component signs must agree exactly under both conventions."""
import numpy as np
import pytest
import torch

from oracle import gwpca_oracle as G

pytestmark = pytest.mark.gpu


def _scene(seed, h, w, c, rank=6, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(h, w, rank)) @ (rng.normal(size=(rank, c)) * np.linspace(3.0, 0.5, rank)[:, None])
    x = np.round((x + 0.2 * rng.normal(size=(h, w, c))) * 400.0 + 6000.0)
    return x.astype(dtype)


def _close(a, b, tol=1e-6):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
    return np.abs(a - b).max() <= tol * np.abs(b).max()


def test_matches_reference_fixture(golden):
    from hsimae_b200.gwpca import applyGWPCA
    z = golden("gwpca.npz")
    major, minor = (int(v) for v in str(z["sklearn_version"]).split(".")[:2])
    sign = "v" if (major, minor) >= (1, 5) else "u"
    for name in ("a", "b"):
        out = applyGWPCA(z[f"{name}/X"], nc=32, group=4, whiten=True, sign=sign)
        assert out.dtype == torch.float64 and tuple(out.shape) == z[f"{name}/whiten"].shape
        assert _close(out, z[f"{name}/whiten"])
        assert _close(applyGWPCA(z[f"{name}/X"], nc=32, group=4, whiten=False, sign=sign), z[f"{name}/plain"])


@pytest.mark.parametrize("h,w,c,group,dtype", [(33, 47, 204, 4, np.float64), (25, 31, 103, 4, np.float32), (40, 40, 120, 2, np.int16),
                                                (19, 23, 224, 8, np.float64), (3, 5, 32, 4, np.float64)])
@pytest.mark.parametrize("whiten", [True, False])
@pytest.mark.parametrize("sign", ["u", "v"])
def test_matches_oracle(h, w, c, group, dtype, whiten, sign):
    from hsimae_b200.gwpca import applyGWPCA
    X = _scene(h * 1000 + c, h, w, c, dtype=dtype)
    nc = 32 if group != 2 else 16
    if h * w < 20:
        nc = 8                                                  # tiny scene: fewer components than pixels
    ref = G.apply_gwpca(X.astype(np.float64), nc, group, whiten, sign)
    out = applyGWPCA(X, nc=nc, group=group, whiten=whiten, sign=sign)
    assert tuple(out.shape) == ref.shape
    assert _close(out, ref), float(np.abs(out.cpu().numpy() - ref).max())
    again = applyGWPCA(torch.from_numpy(X.astype(np.float64) if dtype == np.int16 else X).cuda(), nc=nc, group=group, whiten=whiten, sign=sign)
    assert torch.equal(out, again)                               # fixed reduction order: bit-identical reruns


def test_full_scene_properties():
    """Salinas-sized scene (512 x 217 x 204): whitened group outputs have zero mean, unit variance and are uncorrelated"""
    from hsimae_b200.gwpca import applyGWPCA
    X = torch.from_numpy(_scene(1, 128, 217, 204)).cuda().repeat(4, 1, 1)
    X = X + torch.randn(X.shape, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 40.0
    out = applyGWPCA(X, nc=32, group=4, whiten=True, sign="u").view(-1, 32)
    n = out.shape[0]
    assert n == 512 * 217
    assert float(out.mean(0).abs().max()) < 1e-9
    for g in range(4):
        blk = out[:, 8 * g:8 * g + 8]
        cov = blk.T @ blk / (n - 1)
        assert float((cov - torch.eye(8, dtype=torch.float64, device="cuda")).abs().max()) < 1e-8
    # u-based sign: the largest-magnitude entry of every column is positive
    idx = out.abs().argmax(0)
    assert bool((out[idx, torch.arange(32, device="cuda")] > 0).all())


def test_errors():
    from hsimae_b200.gwpca import applyGWPCA
    with pytest.raises(ValueError):
        applyGWPCA(np.ones((8, 8, 40)))                         # constant scene
    with pytest.raises(ValueError):
        applyGWPCA(_scene(0, 8, 8, 20), nc=32, group=4)         # 5-band groups cannot give 8 components
    with pytest.raises(ValueError):
        applyGWPCA(_scene(0, 8, 8, 300), nc=32, group=4)        # 75-band groups: beyond the 64-band kernel tile
