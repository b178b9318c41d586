#!/bin/bash
# GWPCA on the GPU box: parity tests, Salinas-sized timing, memcheck of the new kernels.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-700; }
run gwpca_tests python -m pytest tests/test_gwpca_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -k "gwpca or packed_weights or loss_curve"
run gwpca_bench python tools/gwpca_bench.py
run gwpca_sanitizer compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gwpca_gpu.py -m gpu -q --tb=short -x -k "fixture or errors or (3-5-32 and True and u) or (19-23-224 and False and u)"
