#!/bin/bash
# round 2, call P: final validation with the final build
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-500; }
TAILN=5 run r2p_tests python -m pytest tests -m gpu -q --tb=short
grep -E "^E  |FAILED" gpurun_out/r2p_tests.log | head
run r2p_bench python bench.py --steps 20 --warmup 5
run r2p_ref python bench.py --impl reference --steps 5 --warmup 2
run r2p_scene python bench.py --workload scene --steps 12
run r2p_ft python bench.py --workload finetune --steps 40 --warmup 5
