#!/bin/bash
# round 2, final validation of the LayerNorm-backward fusion build: tests, bench (+ reference arm), secondary workloads, ncu launch list
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TAILN=5 run r2aa_tests python -m pytest tests -m gpu -q --tb=short
grep -E "^E  |FAILED" gpurun_out/r2aa_tests.log | head
run r2aa_bench python bench.py --steps 20 --warmup 5
run r2aa_ref python bench.py --impl reference --steps 5 --warmup 2
run r2aa_scene python bench.py --workload scene --steps 12
run r2aa_ft python bench.py --workload finetune --steps 40 --warmup 5
HSIMAE_LNBWD_FUSE=0 run r2aa_ft_unfused python bench.py --workload finetune --steps 40 --warmup 5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aa_smoke.log 2>&1; tail -2 gpurun_out/r2aa_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02i_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile > gpurun_out/r2aa_ncu_list.log 2>&1
tail -1 gpurun_out/r2aa_ncu_list.log | cut -c1-200
