#!/bin/bash
echo "default"; python tools/attn_bench.py 2>&1 | tail -1
for v in "-DHSIMAE_ATTN_UNROLL_FWD=2" "-DHSIMAE_ATTN_UNROLL_BWD=2" "-DHSIMAE_ATTN_UNROLL_FWD=2 -DHSIMAE_ATTN_FWD_PER_SM=4"; do
  HSIMAE_NVCC_EXTRA="$v" python -m hsimae_b200.build --force > /dev/null 2>&1
  echo "$v"; HSIMAE_NVCC_EXTRA="$v" python tools/attn_bench.py 2>&1 | tail -1
done
