#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log; }
for fw in 4 8; do
  export HSIMAE_FUSED_MLP_FW=$fw
  TAILN=3 run r2h_mlp_fw$fw python -m pytest tests/test_ops_gpu.py -q --tb=short -k mlp_fused -x
  TAILN=3 run r2h_bench_fw$fw python tools/mlp_bench.py
done
unset HSIMAE_FUSED_MLP_FW
run r2h_model python -m pytest tests/test_model_gpu.py -q --tb=short -x
run r2h_bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline
HSIMAE_FUSED_MLP_FW=8 run r2h_bench8 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
HSIMAE_FUSED_MLP=0 run r2h_bench_off python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
