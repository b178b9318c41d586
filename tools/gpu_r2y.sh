#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_bench.py 2>&1 | tail -2
python tools/lnbwd_bench.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2y_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['loss']); ka=d['kernel_accounting']; print(ka['kernel_time_sum_ms'], {k: round(v['ms'],2) for k,v in ka['families'].items()}, ka['gemm_family_frac']['frac'])"
