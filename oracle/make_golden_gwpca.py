"""Generates tests/golden/gwpca.npz by running the REFERENCE's applyGWPCA (/root/reference/Utils/GroupWisePCA.py) in the
build container (scikit-learn version recorded in the fixture; >= 1.5 fixes component signs v-based).
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_gwpca.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from Utils.GroupWisePCA import applyGWPCA, split_data  # noqa: E402
import sklearn  # noqa: E402


def scene(seed, h, w, c, rank=7):
    """low-rank spectra + noise, scaled like 16-bit sensor counts (distinct leading eigenvalues, noisy tail)"""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(h, w, rank)) @ (rng.normal(size=(rank, c)) * np.linspace(3.0, 0.5, rank)[:, None])
    return np.round((x + 0.2 * rng.normal(size=(h, w, c))) * 400.0 + 6000.0)


out = {"sklearn_version": np.array(sklearn.__version__)}
for name, (h, w, c) in {"a": (12, 11, 52), "b": (17, 9, 103)}.items():
    X = scene(len(name) + c, h, w, c)
    out[f"{name}/X"] = X
    out[f"{name}/whiten"] = applyGWPCA(X, nc=32, group=4, whiten=True)
    out[f"{name}/plain"] = applyGWPCA(X, nc=32, group=4, whiten=False)
    out[f"{name}/widths"] = np.array([p.shape[1] for p in split_data([X.reshape(-1, c)], 4)])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gwpca.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
