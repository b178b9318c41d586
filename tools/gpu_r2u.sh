#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -q -x -k "attention" 2>&1 | tail -2
python tools/attn_bench.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['loss'])"
