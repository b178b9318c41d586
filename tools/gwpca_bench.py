"""Group-wise PCA of a Salinas-sized raw scene (512 x 217 x 204 float64): device path vs the numpy oracle on the host.
Prints one JSON line.  `python tools/gwpca_bench.py`"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsimae_b200.gwpca import applyGWPCA  # noqa: E402
from oracle import gwpca_oracle as G  # noqa: E402

rng = np.random.default_rng(0)
H, W, Cb = 512, 217, 204
X = rng.normal(size=(H, W, 8)) @ rng.normal(size=(8, Cb)) * 300 + 40 * rng.normal(size=(H, W, Cb)) + 5000
xd = torch.from_numpy(X).cuda()
for _ in range(3):
    out = applyGWPCA(xd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    out = applyGWPCA(xd)
torch.cuda.synchronize()
t_dev = (time.perf_counter() - t0) / 10
t0 = time.perf_counter()
for _ in range(10):
    out = applyGWPCA(X)
torch.cuda.synchronize()
t_h2d = (time.perf_counter() - t0) / 10
t0 = time.perf_counter()
ref = G.apply_gwpca(X, sign="u" if out is None else "v")
t_cpu = time.perf_counter() - t0
from hsimae_b200.gwpca import _auto_sign  # noqa: E402
ref = G.apply_gwpca(X, sign=_auto_sign())
err = float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max())
nbytes = 3 * X.nbytes + out.numel() * 8
print(json.dumps({"scene": [H, W, Cb], "device_resident_ms": t_dev * 1e3, "from_host_ms": t_h2d * 1e3, "oracle_cpu_ms": t_cpu * 1e3,
                  "cpu_threads": torch.get_num_threads(), "algorithmic_GBps_device_resident": nbytes / t_dev / 1e9, "max_rel_err_vs_oracle": err}))
