"""Warp-stall hot spots per kernel from an `ncu --set full --import-source on` report (runs without a GPU).
For every kernel: stall-reason totals of the sampled warps and the ten SASS instructions that collected the most samples.
usage: python tools/stall_hotspots.py gpurun_out/prof_kernels.ncu-rep > profiles/rNNx_stall_hotspots.md"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ids = [(r[hdr.index("ID")], r[hdr.index("Kernel Name")], r[hdr.index("gpu__time_duration.sum")]) for r in rows[2:]]
print(f"# Warp-stall hot spots ({rep}; sampled warps, SASS level)\n")
for kid, kname, dur in ids:
    if "hsimae" not in kname:
        continue
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(kid) + 1}"], capture_output=True, text=True).stdout
    lines = src.splitlines()
    start = next((i for i, l in enumerate(lines) if l.startswith('"Address"')), None)
    if start is None:
        continue
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))   # first (SASS) view only
    rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
    stalls = [c for c in rd[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r["# Samples"] or 0) for r in rd) or 1
    agg = {c: sum(int(r[c] or 0) for r in rd) for c in stalls}
    short = re.sub(r"\(.*", "", kname.replace("(int)", "")).replace("void ", "").replace("hsimae::", "").replace("<unnamed>::", "")
    print(f"## `{short}` — {float(dur.replace(',', '')) / 1e3 if float(dur.replace(',', '')) > 1e4 else float(dur.replace(',', '')):.1f} us, {tot} samples\n")
    print("stall reasons: " + ", ".join(f"{c[6:]} {100 * v / tot:.0f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])[:7] if v) + "\n")
    print("| samples | share | instruction | top reason |\n|---|---|---|---|")
    for r in sorted(rd, key=lambda r: -int(r["# Samples"] or 0))[:10]:
        n = int(r["# Samples"] or 0)
        if not n:
            break
        why = max(stalls, key=lambda c: int(r[c] or 0))
        print(f"| {n} | {100 * n / tot:.1f}% | `{r['Source'].strip()[:80]}` | {why[6:]} |")
    print()
