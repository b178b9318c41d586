#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline > gpurun_out/bench_n2.log 2>&1
tail -1 gpurun_out/bench_n2.log | cut -c1-700
