#!/bin/bash
# round 2, call I: graph step parity, configs[3]/[4] workloads, bench with kernel accounting, ncu evidence (summarised on the box:
# the .ncu-rep files are too large to travel back)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-600; }
TAILN=30 run r2i_graph python -m pytest tests/test_graph_gpu.py -q --tb=short -x
grep -E "^E  " gpurun_out/r2i_graph.log | head -20
if [ -z "$SKIP_WL" ]; then
run r2i_ft_graph python bench.py --workload finetune --steps 40 --warmup 5
run r2i_ft_eager python bench.py --workload finetune --steps 40 --warmup 5 --no-graph --no-cpu-baseline
run r2i_scene python bench.py --workload scene --steps 12
fi
run r2i_bench python bench.py --steps 20 --warmup 5
run r2i_ncu_membound ncu --set full --clock-control none --profile-from-start off -k regex:"embed_|loss_kernel|fill_|mask_kernel|ln_bwd|mlp_fused|pack_kernel" -c 40 -o /tmp/r02_membound -f python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
python tools/ncu_step_summary.py /tmp/r02_membound.ncu-rep gpurun_out/r02_ncu_membound.json | tee gpurun_out/r02_ncu_membound.txt
run r2i_ncu_kernels ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_prof_kernels -f python tools/prof_kernels.py
python tools/ncu_summary.py /tmp/r02_prof_kernels.ncu-rep gpurun_out/r02_ncu_full_kernels.json | tee gpurun_out/r02_ncu_full_kernels.txt
run r2i_ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
du -sh gpurun_out
