#!/bin/bash
# full gpu test-suite, first bench line, ncu launch list + full captures of the block kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
run tests python -m pytest tests -m gpu -q --tb=short -x
run bench python bench.py --steps 20 --warmup 5
tail -2 gpurun_out/bench.log | head -1 > gpurun_out/bench_r01_n1.json
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
run ncu_fwd ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"hsimae" -c 12 -o gpurun_out/prof_fwd \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
run ncu_bwd ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"hsimae" -s 290 -c 16 -o gpurun_out/prof_bwd \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
ls -la gpurun_out
