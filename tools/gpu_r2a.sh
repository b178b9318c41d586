#!/bin/bash
# round 2, call A: validate the round-2 groundwork on a B200 (full GPU suite, both bench arms)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -6 gpurun_out/$name.log; }
run r2a_tests python -m pytest tests -m gpu -q --tb=short -x
grep -E "^E  |FAILED|Error" gpurun_out/r2a_tests.log | cut -c1-300 | head -40
run r2a_bench python bench.py --steps 20 --warmup 5
run r2a_ref python bench.py --impl reference --steps 3 --warmup 1
