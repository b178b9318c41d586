#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 20 --warmup 5 --no-roofline --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2af_bench_n$n.json
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/r2af_bench_n1.json
python -c "
import json
b=None
for n in (1,2,4,8):
    d=json.load(open('gpurun_out/r2af_bench_n%d.json'%n)); b = b or d['value']; print(n, round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['value']), d.get('dp_parity_max_rel'), 'eff %.3f' % (d['value']/(n*b)))"
