#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_switches_gpu.py -m gpu -q --tb=short -k "dp or LNBWD" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2ab_bench_n2.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2ab_bench_n1.json
python -c "
import json
for n in (1,2):
    d=json.load(open('gpurun_out/r2ab_bench_n%d.json'%n)); print(n, d['ms_per_step'], d['value'], d['e2e']['value'], d.get('dp_parity_max_rel'))"
