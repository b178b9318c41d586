#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-700; }
TAILN=8 run r2k_tests python -m pytest tests -m gpu -q --tb=short -x
grep -E "^E  |FAILED" gpurun_out/r2k_tests.log | head
run r2k_bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline
HSIMAE_EMBED_MMA=0 run r2k_bench_off python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e
python - <<'PY'
import json
for f in ("r2k_bench", "r2k_bench_off"):
    l = [x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")]
    if not l: continue
    d = json.loads(l[0]); print(f, d["ms_per_step"], d.get("loss"))
    ka = d.get("kernel_accounting")
    if ka and "families" in ka:
        print(" sum", ka["kernel_time_sum_ms"], {k: round(v["ms"], 3) for k, v in ka["families"].items()})
        print(" gemm frac", ka["gemm_family_frac"]["frac"], ka["gemm_family_frac"]["in_step_ms"])
        for k, v in ka["membound_kernels"].items(): print("  ", k, round(v["us"], 1), "us", round(v["gbs"]), "GB/s", round(v["frac_of_hbm_peak"], 3))
PY
