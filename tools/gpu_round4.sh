#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
run attn python -m pytest tests/test_ops_gpu.py -q --tb=short -k attention
grep -E "^E  |FAILED" gpurun_out/attn.log | cut -c1-300 | head -30
run tests python -m pytest tests -m gpu -q --tb=line
run bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for row in csv.DictReader(lines):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    n=re.sub(r'\(.*','',row['Kernel Name']); agg[n][0]+=1; agg[n][1]+=v; tot+=v
print(f"total {tot/1e3:.2f} ms, {sum(a[0] for a in agg.values())} launches")
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{t:9.0f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:8.1f} {k[:80]}")
PY
