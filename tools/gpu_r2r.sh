#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=5 run r2r_tests python -m pytest tests -m gpu -q --tb=short -x
grep -E "^E  |FAILED" gpurun_out/r2r_tests.log | head
run r2r_bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline
HSIMAE_OVERLAP=0 run r2r_bench_off python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
run r2r_scene python bench.py --workload scene --steps 12 --no-cpu-baseline
HSIMAE_OVERLAP=0 run r2r_scene_off python bench.py --workload scene --steps 12 --no-cpu-baseline
