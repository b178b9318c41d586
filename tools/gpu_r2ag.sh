#!/bin/bash
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']))"; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', round(d['ms_per_step'],3), round(d['value']))"
echo "default"; run 29601
for c in 1 2 4 8; do echo "NCCL_MAX_NCHANNELS=$c"; NCCL_MAX_NCHANNELS=$c run 2961$c; done
for c in 1 4; do echo "NCCL_MAX_CTAS=$c"; NCCL_MAX_CTAS=$c run 2962$c; done
