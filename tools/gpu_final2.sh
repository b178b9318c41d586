#!/bin/bash
# last call of round 2: the driver's own sequence on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>&1 | grep '^{' > gpurun_out/r2_final_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'])"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep '^{' | cut -c1-200
