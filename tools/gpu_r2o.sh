#!/bin/bash
# round 2, call O: evidence with the final build -- launch list, ncu of the backward HBM-bound kernels and the new kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
run r2o_ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
run r2o_ncu_membound ncu --set full --clock-control none --profile-from-start off -k regex:"embed_|loss_kernel|fill_|mask_kernel|ln_bwd|pack_kernel" -c 24 -o /tmp/r02_membound -f python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
python tools/ncu_step_summary.py /tmp/r02_membound.ncu-rep gpurun_out/r02f_ncu_membound.json | tee gpurun_out/r02f_ncu_membound.txt
run r2o_ncu_new ncu --set full --clock-control none --profile-from-start off -k regex:"mlp_fused|wgrad_group|embed_bwd|fill_bwd" -s 20 -c 16 -o /tmp/r02_new -f python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
python tools/ncu_step_summary.py /tmp/r02_new.ncu-rep gpurun_out/r02f_ncu_new_kernels.json | tee gpurun_out/r02f_ncu_new_kernels.txt
run r2o_smoke python -c "import __graft_entry__ as g; g.smoke()"
du -sh gpurun_out
