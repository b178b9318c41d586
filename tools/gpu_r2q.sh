#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do HSIMAE_GEMM_N256=$v python tools/gemm_n256_probe.py 2>&1 | tail -1; done | tee gpurun_out/r2q_n256.log
python -m pytest tests/test_ops_gpu.py -q -x -k "gemm_bias or strided" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e 2>&1 | grep '^{' | cut -c1-220
HSIMAE_GEMM_N256=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e 2>&1 | grep '^{' | cut -c1-220
