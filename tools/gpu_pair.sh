#!/bin/bash
mkdir -p gpurun_out
echo "== ops tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -q --tb=short -x -k "gemm or wgrad" 2>&1 | tail -15
timeout 300 python tools/gemm_bench.py 2>&1 | tail -3
HSIMAE_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py 2>&1 | tail -3
HSIMAE_GEMM_PAIR=2 HSIMAE_GEMM_ARES_N_GATE=256 timeout 300 python tools/gemm_bench.py 2>&1 | tail -3
HSIMAE_GEMM_ARES_N=128 HSIMAE_WGRAD_PAIR=0 timeout 300 python tools/gemm_bench.py 2>&1 | tail -3
echo "== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q --tb=short -x 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
HSIMAE_SAVE_GATE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-roofline 2>&1 | tail -1 | cut -c1-330
