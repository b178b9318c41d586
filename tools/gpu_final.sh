#!/bin/bash
# Round-end evidence: GPU tests, bench line, ncu launch list of one step, ncu --set full of the hot kernels, bandwidth calibration.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests -m gpu -q --tb=line
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run membw python tools/membw.py
run bench python bench.py --steps 20 --warmup 5
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
run prof_kernels ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_kernels python tools/prof_kernels.py
ls -la gpurun_out/*.ncu-rep
