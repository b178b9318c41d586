#!/bin/bash
# round 2, call B: fused MLP kernel parity (op level), then whole-model suites with it on the product path
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run r2b_mlp python -m pytest tests/test_ops_gpu.py -q --tb=short -k mlp_fused -x
grep -E "^E  |FAILED|Error|timeout" gpurun_out/r2b_mlp.log | cut -c1-300 | head -30
if grep -q "passed" gpurun_out/r2b_mlp.log && ! grep -q "failed" gpurun_out/r2b_mlp.log; then
  run r2b_model python -m pytest tests/test_model_gpu.py -q --tb=short -x
  grep -E "^E  |FAILED|Error" gpurun_out/r2b_model.log | cut -c1-300 | head -30
  run r2b_bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline
  HSIMAE_FUSED_MLP=0 run r2b_bench_off python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
fi
