#!/bin/bash
# Short round-end check: GPU tests, smoke, one bench line (about 90 s of box time).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-600; }
run tests python -m pytest tests -m gpu -q --tb=short -s -k "not test_switch_combination"
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench python bench.py --steps 20 --warmup 5
