#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-600; }
TAILN=12 run r2n_wg python -m pytest tests/test_ops_gpu.py -q --tb=short -x -k "wgrad"
grep -E "^E  |FAILED" gpurun_out/r2n_wg.log | head
run r2n_model python -m pytest tests/test_model_gpu.py -q --tb=short -x
grep -E "^E  |FAILED" gpurun_out/r2n_model.log | head
run r2n_bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline
HSIMAE_WGRAD_GROUP=0 run r2n_bench_off python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
python - <<'PY'
import json
for f in ("r2n_bench", "r2n_bench_off"):
    l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l: continue
    d = json.loads(l[0]); print(f, d["ms_per_step"], d.get("loss"), d["gpu_launches_per_step"])
    ka = d.get("kernel_accounting")
    if ka and "families" in ka:
        print(" sum", ka["kernel_time_sum_ms"], {k: round(v["ms"], 3) for k, v in ka["families"].items()})
        print(" gemm frac", ka["gemm_family_frac"]["frac"], ka["gemm_family_frac"]["in_step_ms"])
        for t in ka["top_kernels"][:8]: print("   ", t)
PY
