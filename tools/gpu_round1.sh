#!/bin/bash
# first GPU bring-up: operator parity (tcgen05 vs checker vs torch), then model parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run mask      python -m pytest tests/test_ops_gpu.py -q --tb=short -k "mask"
run attn      python -m pytest tests/test_ops_gpu.py -q --tb=short -k "attention"
run gemm_bias python -m pytest tests/test_ops_gpu.py -q --tb=short -k "gemm_bias or strided"
run gemm_ln   python -m pytest tests/test_ops_gpu.py -q --tb=short -k "resid_layernorm"
run gemm_glu  python -m pytest tests/test_ops_gpu.py -q --tb=short -k "swiglu"
run wgrad     python -m pytest tests/test_ops_gpu.py -q --tb=short -k "wgrad"
HSIMAE_DEBUG_SIMT=1 run model_simt python -m pytest tests/test_model_gpu.py -q --tb=short -k "not full_size"
run model_tc  python -m pytest tests/test_model_gpu.py -q --tb=short
run smoke     python __graft_entry__.py smoke
