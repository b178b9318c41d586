import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the live mount in the build container; on the GPU box the byte-identical copy made by oracle/fetch_ref.py
REFERENCE = "/root/reference"
if not os.path.exists(os.path.join(REFERENCE, "Models.py")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Models.py")):
    REFERENCE = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
    return load


def state_from_npz(z, device="cpu"):
    return {k[3:]: torch.from_numpy(v).to(device) for k, v in z.items() if k.startswith("sd/")}


def grads_from_npz(z):
    return {k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("grad/")}


TINY = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=64, depth=3, num_heads=4, s_depth=2,
            decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=4, norm_pix_loss=True, trunc_init=True)


def tiny_geometry(num_class=0):
    from oracle.hsimae_oracle import Geometry
    return Geometry(embed_dim=64, depth=3, s_depth=2, num_heads=4, decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=4,
                    num_class=num_class)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64"""
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
