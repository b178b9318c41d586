#!/bin/bash
# ncu launch list of one training step + per-kernel summary
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for row in csv.DictReader(lines):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    n=re.sub(r'\(.*','',row['Kernel Name']); agg[n][0]+=1; agg[n][1]+=v; tot+=v
print(f"total {tot/1e3:.2f} ms, {sum(a[0] for a in agg.values())} launches")
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:24]:
    print(f"{t:9.0f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:8.1f} {k[:90]}")
PY
