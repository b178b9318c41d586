"""CPU oracle for the group-wise PCA preprocessing.  TEST INFRASTRUCTURE ONLY: imported by tests/ and by
tools/gwpca_bench.py's CPU leg, never by the product path (hsimae_b200/gwpca.py runs on the device).

numpy restatement of applyGWPCA (/root/reference/Utils/GroupWisePCA.py:20-34) with sklearn's PCA (third-party: the
reference pins scikit-learn 1.3.2, this container has 1.9) restated from its published algorithm: centre, eigen-decompose
the covariance (divisor n-1), keep the leading components, fix their signs with svd_flip, project, divide by
sqrt(explained variance) when whitening.  Pinned against the live reference function (sklearn 1.9: v-based signs) and
sklearn.utils.extmath.svd_flip (u-based signs, the convention of sklearn <= 1.4) in tests/test_gwpca_cpu.py, and against
the fixture tests/golden/gwpca.npz generated from the reference by oracle/make_golden_gwpca.py."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def band_groups(c: int, group: int = 4) -> List[Tuple[int, int]]:
    """(offset, width) of the contiguous band groups split_data produces (GroupWisePCA.py:5-17): `group // 2` rounds of
    halving every piece at `c // 2` -- i.e. 2 ** (group // 2) groups (4 for the reference's group=4)."""
    pieces = [(0, c)]
    for _ in range(group // 2):
        nxt = []
        for off, w in pieces:
            nxt += [(off, w // 2), (off + w // 2, w - w // 2)]
        pieces = nxt
    return pieces


def pca_fit_transform(x: np.ndarray, k: int, whiten: bool, sign: str = "v"):
    """sklearn.decomposition.PCA(n_components=k, whiten=whiten).fit_transform(x) (GroupWisePCA.py:28-29) ->
    (transformed [n, k], components [k, b], explained_variance [k])"""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    xc = x - x.mean(axis=0)
    cov = xc.T @ xc / (n - 1)
    lam, vec = np.linalg.eigh(cov)
    order = np.argsort(lam)[::-1][:k]
    lam, comps = np.maximum(lam[order], 0.0), vec[:, order].T.copy()
    proj = xc @ comps.T
    if sign == "v":      # svd_flip(u, v, u_based_decision=False): largest-magnitude entry of every component positive
        s = np.sign(comps[np.arange(k), np.argmax(np.abs(comps), axis=1)])
    elif sign == "u":    # svd_flip(u, v): largest-magnitude entry of every column of u (first on ties) positive
        s = np.sign(proj[np.argmax(np.abs(proj), axis=0), np.arange(k)])
    else:
        raise ValueError(sign)
    s[s == 0] = 1.0
    comps *= s[:, None]
    proj *= s[None, :]
    if whiten:
        proj = proj / np.sqrt(lam)[None, :]
    return proj, comps, lam


def apply_gwpca(X: np.ndarray, nc: int = 32, group: int = 4, whiten: bool = True, sign: str = "v") -> np.ndarray:
    """applyGWPCA (GroupWisePCA.py:20-34): [h, w, c] -> [h, w, n_groups * (nc // group)] float64"""
    h, w, c = X.shape
    X = np.reshape(X, (-1, c)).astype(np.float64)
    X = (X - X.min()) / (np.max(X) - np.min(X))
    outs = [pca_fit_transform(X[:, off:off + b], nc // group, whiten, sign)[0] for off, b in band_groups(c, group)]
    return np.concatenate(outs, axis=-1).reshape(h, w, -1)
