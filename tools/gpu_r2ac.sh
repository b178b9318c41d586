#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --tb=short -k "lnbwd and not 40000" > gpurun_out/r2ac_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/r2ac_memcheck.log | cut -c1-300
HSIMAE_LNBWD_MIN_ROWS=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_model_gpu.py -m gpu -q -x --tb=short -k "reference_configs and 256-16-40" > gpurun_out/r2ac_memcheck_model.log 2>&1; echo "memcheck(model) exit $?"; tail -4 gpurun_out/r2ac_memcheck_model.log | cut -c1-300
