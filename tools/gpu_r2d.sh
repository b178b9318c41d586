#!/bin/bash
mkdir -p gpurun_out
HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE python -c "from hsimae_b200 import build; build.build(force=True)" 2>&1 | tail -3
timeout 300 python tools/mlp_trace.py 2>&1 | tee gpurun_out/r2d_mlp_trace.log | tail -40
