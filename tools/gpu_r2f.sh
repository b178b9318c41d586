#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 4 6; do echo "== DBG=$v"; HSIMAE_FUSED_MLP_DBG=$v timeout 300 python tools/mlp_bench.py 2>&1 | grep encoder | cut -c1-200; done | tee gpurun_out/r2f_dbg.log
