#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -q -x -k "mlp_fused" 2>&1 | tail -2
HSIMAE_FUSED_MLP_FW=8 python -m pytest tests/test_ops_gpu.py -q -x -k "mlp_fused" 2>&1 | tail -1
python tools/mlp_bench.py 2>&1 | tail -2 | cut -c1-220
python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'])"
