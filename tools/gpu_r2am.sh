#!/bin/bash
HSIMAE_NVCC_EXTRA=-DHSIMAE_TRACE python -m hsimae_b200.build --force > /dev/null 2>&1
python tools/mlp_trace.py 2>&1 | tail -30 | cut -c1-250
