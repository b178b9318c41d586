#!/bin/bash
mkdir -p gpurun_out
echo "== ops tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -q --tb=short -x 2>&1 | tail -5
timeout 300 python tools/gemm_bench.py 2>&1 | tail -3
HSIMAE_GEMM_PAIR=2 timeout 300 python tools/gemm_bench.py 2>&1 | tail -3 | head -1
HSIMAE_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py 2>&1 | tail -3 | head -1
echo "== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q --tb=short -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
