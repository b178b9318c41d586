#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ft_launches.csv python bench.py --workload finetune --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2l_ft.log 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/r02_ft_launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
# last step only: take the final third of launches
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
n=len(rows); rows=rows[-(n//7):]
for row in rows:
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    k=re.sub(r'\(.*','',row['Kernel Name']).replace('void ',''); agg[k][0]+=1; agg[k][1]+=v; tot+=v
print(f"last step: total {tot/1e3:.2f} ms, {sum(a[0] for a in agg.values())} launches")
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:22]:
    print(f"{t:9.0f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:8.1f} {k[:90]}")
PY
