"""Turn an ncu launch list (CSV of gpu__time_duration.sum) into the per-kernel markdown table kept under profiles/.
usage: python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/rNNx_launch_summary.md"""
import csv, collections, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    agg[k][0] += 1; agg[k][1] += v; tot += v; n += 1
print(f"# {sys.argv[2]}\n")
print("command: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --profile`\n")
print(f"total {tot / 1e3:.2f} ms over {n} launches (serialised, cold-cache per-launch times; the kernel SHARES are what carries over to the un-profiled step)\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:26]:
    print(f"| `{k[:90]}` | {c} | {t:.0f} | {100 * t / tot:.1f}% | {t / c:.1f} |")
