"""Group-wise PCA preprocessing on the device.

Host-side mirror of the reference's `applyGWPCA` (/root/reference/Utils/GroupWisePCA.py:20-34; called by
`get_data_cut_file` / `get_data_set_dual`, Utils/Preprocessing.py:90-91,192-193): same name, arguments and result
layout.  The scene is uploaded once; two kernels sweep the pixels for the global extrema, the band means and the
covariance of every band group (`hsimae_gwpca_moments`), the host solves the tiny (<= 64 x 64) symmetric eigenproblems,
and one kernel projects (`hsimae_gwpca_project`).  There is no CPU path.

Sign convention of the components (PCA is defined up to a sign per component; sklearn fixes it with `svd_flip`):
  sign="u"  largest-magnitude entry of every transformed column positive  (scikit-learn <= 1.4; the reference pins 1.3.2)
  sign="v"  largest-magnitude entry of every component vector positive    (scikit-learn >= 1.5)
  sign="auto" (default) follows the scikit-learn installed next to the reference, "u" if there is none.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _lib


def band_groups(c: int, group: int = 4) -> List[Tuple[int, int]]:
    """(offset, width) per band group: `group // 2` rounds of halving at `c // 2` (split_data, GroupWisePCA.py:5-17)."""
    pieces = [(0, c)]
    for _ in range(group // 2):
        pieces = [q for off, w in pieces for q in ((off, w // 2), (off + w // 2, w - w // 2))]
    return pieces


def _auto_sign() -> str:
    try:
        import sklearn
        major, minor = (int(v) for v in sklearn.__version__.split(".")[:2])
        return "v" if (major, minor) >= (1, 5) else "u"
    except Exception:
        return "u"


def applyGWPCA(X, nc: int = 32, group: int = 4, whiten: bool = True, sign: str = "auto", device="cuda:0") -> torch.Tensor:
    """[h, w, c] scene (numpy array or tensor, any real dtype) -> float64 tensor [h, w, n_groups * (nc // group)] on `device`."""
    if sign == "auto":
        sign = _auto_sign()
    if sign not in ("u", "v"):
        raise ValueError("sign must be 'u', 'v' or 'auto'")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hsimae_b200.gwpca.applyGWPCA runs on a CUDA device only; there is no CPU path")
    t = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X))
    if t.dim() != 3:
        raise ValueError("X must be [h, w, c]")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)                              # integer sensors counts: exact in fp64, as numpy promotes them
    h, w, c = t.shape
    n = h * w
    groups = band_groups(c, group)
    k = nc // group
    widths = [b for _, b in groups]
    if k < 1 or min(widths) < k or n < 2:
        raise ValueError(f"n_components={k} must be between 1 and min(n_samples, n_features)={min(min(widths), n)}")   # as sklearn raises
    if max(widths) > 64 or len(groups) > 16 or len(groups) * k > 64:
        raise ValueError("hsimae_b200.gwpca: at most 16 groups of at most 64 bands and 64 output channels are supported")
    x = t.to(dev).contiguous().view(n, c)
    L = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ng, nout = len(groups), len(groups) * k
    off = (C.c_int32 * (ng + 1))(*([o for o, _ in groups] + [c]))
    dt = 0 if x.dtype == torch.float32 else 1
    ws = torch.empty(L.hsimae_gwpca_workspace_bytes(c, ng, nout), dtype=torch.uint8, device=dev)
    mean = torch.empty(c, dtype=torch.float64, device=dev)
    minmax = torch.empty(2, dtype=torch.float64, device=dev)
    cov = torch.empty(ng, 64, 64, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):                    # the launches go to the current device's context
        _lib.check(L.hsimae_gwpca_moments(x.data_ptr(), dt, n, c, ng, off, ws.data_ptr(), ws.numel(), mean.data_ptr(), minmax.data_ptr(),
                                          cov.data_ptr(), st), "gwpca_moments")
    cov_h, (lo, hi) = cov.cpu().numpy(), minmax.cpu().tolist()
    rng = hi - lo
    if not rng > 0:
        raise ValueError("constant scene: max == min (division by zero in the reference as well)")
    Wm = np.zeros((nout, 64), dtype=np.float64)
    for g, (_, b) in enumerate(groups):
        lam, vec = np.linalg.eigh(cov_h[g, :b, :b])          # raw-value covariance = range^2 x the normalised one
        order = np.argsort(lam)[::-1][:k]
        lam, comps = np.maximum(lam[order], 0.0), vec[:, order].T
        if sign == "v":
            s = np.sign(comps[np.arange(k), np.argmax(np.abs(comps), axis=1)])
            s[s == 0] = 1.0
            comps = comps * s[:, None]
        # (x' - mean') v = (x - mean) v / range and explained_variance' = lam / range^2: the range cancels when whitening
        scale = 1.0 / np.sqrt(lam) if whiten else np.full(k, 1.0 / rng)
        Wm[g * k:(g + 1) * k, :b] = comps * scale[:, None]
    W = torch.from_numpy(Wm).to(dev)
    out = torch.empty(n, nout, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.hsimae_gwpca_project(x.data_ptr(), dt, n, c, ng, off, k, mean.data_ptr(), W.data_ptr(), out.data_ptr(),
                                          int(sign == "u"), ws.data_ptr(), ws.numel(), st), "gwpca_project")
    return out.view(h, w, nout)
