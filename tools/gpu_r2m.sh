#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-900; }
TAILN=6 run r2m_tests python -m pytest tests -m gpu -q --tb=short -x
grep -E "^E  |FAILED" gpurun_out/r2m_tests.log | head
run r2m_ft_graph python bench.py --workload finetune --steps 40 --warmup 5 --no-cpu-baseline
run r2m_ft_graph_fused python bench.py --workload finetune --steps 40 --warmup 5 --no-cpu-baseline --fused-optimizer
run r2m_ft_eager python bench.py --workload finetune --steps 40 --warmup 5 --no-graph --no-cpu-baseline
run r2m_scene python bench.py --workload scene --steps 12 --no-cpu-baseline
