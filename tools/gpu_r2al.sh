#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "attention" > gpurun_out/r2al_memcheck_attn.log 2>&1; echo "memcheck(attention) exit $?"; tail -4 gpurun_out/r2al_memcheck_attn.log | cut -c1-200
HSIMAE_ATTN_SMALL=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "attention" > gpurun_out/r2al_memcheck_attn_mma.log 2>&1; echo "memcheck(attention, mma forced) exit $?"; tail -3 gpurun_out/r2al_memcheck_attn_mma.log | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "attention and (37-256-16-3-6-full or 300-64-8-4-9-full or 37-256-16-3-6-spatial)" > gpurun_out/r2al_racecheck_attn.log 2>&1; echo "racecheck(attention) exit $?"; tail -4 gpurun_out/r2al_racecheck_attn.log | cut -c1-200
