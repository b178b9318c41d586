#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
run r2e_mlp python -m pytest tests/test_ops_gpu.py -q --tb=short -k mlp_fused -x
grep -E "^E  |FAILED|Error|timeout" gpurun_out/r2e_mlp.log | cut -c1-300 | head -20
TAILN=3 run r2e_mlp_bench python tools/mlp_bench.py
HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE python -c "from hsimae_b200 import build; build.build(force=True)" 2>&1 | tail -3
TAILN=30 run r2e_trace python tools/mlp_trace.py
