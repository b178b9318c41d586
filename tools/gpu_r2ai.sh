#!/bin/bash
# round 2, last call: validation of the final build + evidence (tests, bench, reference arm, smoke, ncu launch list, ncu --set full of the hot kernels)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=5 run r2ai_tests python -m pytest tests -m gpu -q --tb=short
grep -E "^E  |FAILED" gpurun_out/r2ai_tests.log | head
run r2ai_bench python bench.py --steps 20 --warmup 5
run r2ai_ref python bench.py --impl reference --steps 5 --warmup 2
run r2ai_scene python bench.py --workload scene --steps 12
run r2ai_ft python bench.py --workload finetune --steps 40 --warmup 5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai_smoke.log 2>&1; tail -2 gpurun_out/r2ai_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02j_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile > gpurun_out/r2ai_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels_r02j python tools/prof_kernels.py > gpurun_out/prof_kernels_r02j.log 2>&1
ls -la gpurun_out/prof_kernels_r02j.ncu-rep
