#!/bin/bash
mkdir -p gpurun_out
for r in 0 1 2 4 7 3; do
HSIMAE_REVERSE=$r python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-eager 2>&1 | grep '^{' > gpurun_out/r2ak_bench_$r.json; python -c "
import json; d=json.load(open('gpurun_out/r2ak_bench_$r.json')); ka=d['kernel_accounting']; print('rev=$r', round(d['ms_per_step'],3), round(d['value']), round(d['loss'],6), round(ka['kernel_time_sum_ms'],2), {k[:12]: round(v['ms'],2) for k,v in ka['families'].items() if v['ms']>0.5})"
done
HSIMAE_REVERSE=7 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -3
