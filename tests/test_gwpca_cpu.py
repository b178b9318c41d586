"""CPU: the group-wise PCA oracle (oracle/gwpca_oracle.py) against the fixture the reference produced, the live reference
function and sklearn's own sign rule; host-side logic of the device mirror (hsimae_b200/gwpca.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import REFERENCE
from oracle import gwpca_oracle as G


def _close(a, b, tol=1e-6):
    return np.abs(a - b).max() <= tol * np.abs(b).max()


def test_oracle_matches_reference_fixture(golden):
    z = golden("gwpca.npz")
    major, minor = (int(v) for v in str(z["sklearn_version"]).split(".")[:2])
    sign = "v" if (major, minor) >= (1, 5) else "u"
    for name in ("a", "b"):
        X = z[f"{name}/X"]
        assert [b for _, b in G.band_groups(X.shape[2], 4)] == list(z[f"{name}/widths"])
        assert _close(G.apply_gwpca(X, 32, 4, True, sign), z[f"{name}/whiten"])
        assert _close(G.apply_gwpca(X, 32, 4, False, sign), z[f"{name}/plain"])


def test_band_groups_quirks():
    # `group // 2` halving rounds => 2 ** (group // 2) groups, uneven halves go low-first (GroupWisePCA.py:5-17)
    assert G.band_groups(204, 4) == [(0, 51), (51, 51), (102, 51), (153, 51)]
    assert G.band_groups(103, 4) == [(0, 25), (25, 26), (51, 26), (77, 26)]
    assert len(G.band_groups(204, 8)) == 16 and len(G.band_groups(204, 2)) == 2
    from hsimae_b200 import gwpca
    for c in (32, 103, 144, 200, 204, 224):
        for g in (2, 4, 8):
            assert gwpca.band_groups(c, g) == G.band_groups(c, g)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_oracle_matches_live_reference():
    sys.path.insert(0, REFERENCE)
    try:
        from Utils.GroupWisePCA import applyGWPCA
    finally:
        sys.path.remove(REFERENCE)
    import sklearn
    major, minor = (int(v) for v in sklearn.__version__.split(".")[:2])
    sign = "v" if (major, minor) >= (1, 5) else "u"
    rng = np.random.default_rng(5)
    for h, w, c, group in ((20, 30, 204, 4), (16, 16, 144, 2), (30, 30, 200, 4)):
        X = rng.normal(size=(h, w, 5)) @ rng.normal(size=(5, c)) * 300 + 30 * rng.normal(size=(h, w, c)) + 2000
        for whiten in (True, False):
            assert _close(G.apply_gwpca(X, 32, group, whiten, sign), applyGWPCA(X, nc=32, group=group, whiten=whiten))


def test_u_sign_rule_is_sklearns_svd_flip():
    extmath = pytest.importorskip("sklearn.utils.extmath")
    rng = np.random.default_rng(9)
    x = rng.normal(size=(300, 20)) * np.linspace(5, 0.5, 20)
    xc = x - x.mean(0)
    U, S, Vt = np.linalg.svd(xc, full_matrices=False)
    for mode, based in (("u", True), ("v", False)):
        u, vt = extmath.svd_flip(U.copy(), Vt.copy(), u_based_decision=based)
        proj, comps, lam = G.pca_fit_transform(x, 6, True, mode)
        assert _close(comps, vt[:6]) and _close(proj, u[:, :6] * np.sqrt(299.0)) and _close(lam, S[:6] ** 2 / 299.0)


def test_device_mirror_has_no_cpu_path():
    from hsimae_b200 import gwpca
    with pytest.raises(RuntimeError):
        gwpca.applyGWPCA(np.zeros((4, 4, 40)), device="cpu")
    with pytest.raises(ValueError):
        gwpca.applyGWPCA(np.zeros((4, 4, 40)), sign="w")
