"""Per-kernel SASS mnemonic census of the built library (runs without a GPU): which kernels use the Blackwell tensor /
TMA paths.  `python tools/sass_audit.py > profiles/rNNx_sass_audit.md`
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit,
HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async, RED = red.global.add."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hsimae_b200", "libhsimae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "LDGSTS", "RED", "DFMA", "ELECT", "ACQBULK", "SYNCS"]
census, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); census[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        census[cur]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                census[cur][w] += 1
print("# SASS census of hsimae_b200/libhsimae_b200.so (cuobjdump -sass; sm_100a)\n")
print("| kernel | instr | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH))
for fn, c in census.items():
    name = demangle(fn).replace("(int)", "").replace("(bool)", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("hsimae::", "")
    print(f"| `{name[:70]}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
