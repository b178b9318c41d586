#!/bin/bash
mkdir -p gpurun_out
for v in "" "HSIMAE_FUSED_MLP_STRICT=1" "HSIMAE_FUSED_MLP_PAIR=0"; do
  echo "=== mlp_bench $v"; env $v timeout 300 python tools/mlp_bench.py 2>&1 | tail -3
done | tee gpurun_out/r2c_mlp_bench.log
