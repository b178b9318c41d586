#!/bin/bash
# round 2, call J (2 GPUs): NCCL gradient parity, 2-rank bench with the six-bucket exchange
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-900; }
run r2j_tests python -m pytest tests/test_dp_gpu.py tests/test_model_gpu.py tests/test_graph_gpu.py -q --tb=short -x
grep -E "^E  " gpurun_out/r2j_tests.log | head
run r2j_bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline
run r2j_bench_n1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline
