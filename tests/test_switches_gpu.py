"""GPU: every tuning switch of DESIGN.md section 4c still gives reference-matching results.  The switches are read once
per process, so each combination runs one of the model parity tests in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMBOS = [
    {"HSIMAE_SAVE_GATE": "1"},                                   # saved pre-activations instead of the recompute kernel
    {"HSIMAE_GEMM_PAIR": "0", "HSIMAE_WGRAD_PAIR": "0"},         # no CTA pairs anywhere
    {"HSIMAE_GEMM_PAIR": "2"},                                   # CTA pairs wherever legal
    {"HSIMAE_PDL": "0", "HSIMAE_ATTN_SMALL": "0"},               # plain launches, mma attention for short groups
    {"HSIMAE_GEMM_ARES": "0", "HSIMAE_WGRAD_FUSED_BIAS": "0"},   # streaming kernels only, separate bias-gradient kernel
    {"HSIMAE_GEMM_ARES_N": "128", "HSIMAE_GEMM_ARES_N_GATE": "256"},
    # round 2: the fused gated-MLP kernel with eight final-epilogue warps / in its single-CTA form; the two-launch MLP and
    # the CUDA-core patch embedding; saved pre-activations + per-output-tile kernels below a row threshold
    {"HSIMAE_FUSED_MLP_FW": "8"},
    {"HSIMAE_FUSED_MLP_PAIR": "0"},
    {"HSIMAE_FUSED_MLP": "0", "HSIMAE_EMBED_MMA": "0"},
    {"HSIMAE_RECOMPUTE_MIN_ROWS": "8192"},
    {"HSIMAE_LNBWD_FUSE": "0"},                                  # dgrad GEMM and LayerNorm backward as two launches
    {"HSIMAE_LNBWD_MIN_ROWS": "0"},                              # ... and as one kernel at every batch size (default: from two waves of row tiles)
    {"HSIMAE_LNBWD_MIN_ROWS": "0", "HSIMAE_GEMM_PAIR": "2"},     # ... in its CTA-pair form
    {"HSIMAE_WGRAD_GROUP": "0", "HSIMAE_OVERLAP": "0"},          # ... and both encoder chains on one stream                                 # one weight-gradient launch per problem
]


@pytest.mark.parametrize("env", COMBOS, ids=lambda e: ",".join(f"{k[7:]}={v}" for k, v in e.items()))
def test_switch_combination_keeps_parity(env):
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "--tb=short", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_model_gpu.py"), "-k", "reference_configs and (256-16-40 or 128-8-21)"]
    r = subprocess.run(cmd, cwd=ROOT, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout, r.stdout[-500:]
