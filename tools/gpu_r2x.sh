#!/bin/bash
for d in 0 1 2 4 8 16 32 63; do echo "dbg=$d"; HSIMAE_LNBWD_DBG=$d python tools/lnbwd_bench.py 2>&1 | tail -3; done
