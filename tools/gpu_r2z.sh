#!/bin/bash
mkdir -p gpurun_out
HSIMAE_LNBWD_MIN_N=32 timeout 300 python - <<'PY'
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_ops_gpu as T
from hsimae_b200 import ops
for shape in [(300, 64, 192), (5000, 64, 352), (1000, 32, 64), (1000, 128, 384), (147456, 64, 192)]:
    for v in ("plain", "inplace+scale", "no_dxb"):
        T._lnbwd_case(ops, *shape, v)
print("narrow lnbwd cases ok")
PY
for n in 32 160; do
HSIMAE_LNBWD_MIN_N=$n python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2z_bench_$n.json; python -c "
import json; d=json.load(open('gpurun_out/r2z_bench_$n.json')); print($n, d['ms_per_step'], d['value'], d['e2e']['value'], d['loss']); ka=d['kernel_accounting']; print(ka['kernel_time_sum_ms'], {k: round(v['ms'],2) for k,v in ka['families'].items()}, ka['gemm_family_frac']['frac'], ka['gemm_family_frac']['frac_excl_fused_ln']); print(d['roofline']['kernel'], d['roofline']['frac'], {k[:40]: round(v['us'],1) for k,v in d['roofline']['all_kernels'].items()})"
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels_r02i python tools/prof_kernels.py > gpurun_out/prof_kernels_r02i.log 2>&1
tail -2 gpurun_out/prof_kernels_r02i.log
ls -la gpurun_out/*.ncu-rep
