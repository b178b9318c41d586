#!/bin/bash
# GEMM bring-up: operator tests, then A/B timings of the tuning switches.
mkdir -p gpurun_out
echo "== ops tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -q --tb=short -x -k "gemm or wgrad" 2>&1 | tail -15
for w in 1 0; do
  HSIMAE_WGRAD_PAIR=$w timeout 300 python tools/gemm_bench.py 2>&1 | tail -2
done
