#!/bin/bash
# Round-end evidence (about 3 min of box time): GPU tests, smoke, bench line + reference arm, ncu launch list of one step,
# the two secondary configurations (fine-tuning step, dense scene classification).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests -m gpu -q --tb=short
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench python bench.py --steps 20 --warmup 5
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
run finetune python tools/bench_finetune.py
run scene python tools/bench_scene.py
