#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "lnbwd" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -5
for f in 1 0; do
HSIMAE_LNBWD_FUSE=$f python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2w_bench_$f.json; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_$f.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['loss']); ka=d['kernel_accounting']; print(ka['kernel_time_sum_ms'], {k: round(v['ms'],2) for k,v in ka['families'].items()}, ka['gemm_family_frac']['frac'])"
done
