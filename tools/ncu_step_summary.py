"""Summarise an `ncu --set full` capture of kernels of a real bench step into JSON: one entry per kernel NAME (launches averaged)
with duration, DRAM bytes read / written per launch, achieved DRAM GB/s against the measured HBM peak, tensor-pipe activity.

usage: python tools/ncu_step_summary.py gpurun_out/r02_membound.ncu-rep profiles/r02_ncu_membound.json"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def val(r, name):
        if name not in col:
            return None
        try:
            v = float(r[col[name]].replace(",", ""))
        except ValueError:
            return None
        u = units[col[name]]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)

    agg = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
        a = agg.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "tensor": [], "issue": [],
                                  "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "regs": val(r, "launch__registers_per_thread"),
                                  "smem_dyn_bytes": val(r, "launch__shared_mem_per_block_dynamic")})
        a["launches"] += 1
        a["time_us"] += val(r, "gpu__time_duration.sum") or 0.0
        a["dram_read_bytes"] += val(r, "dram__bytes_read.sum") or 0.0
        a["dram_write_bytes"] += val(r, "dram__bytes_write.sum") or 0.0
        for k in col:
            if k == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" or k.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                v = val(r, k)
                if v is not None:
                    a["tensor"].append(v)
        v = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        if v is not None:
            a["issue"].append(v)
    res = []
    for name, a in agg.items():
        n = a["launches"]
        t = a["time_us"] / n
        e = {"kernel": name, "launches": n, "time_us": t, "dram_read_bytes": a["dram_read_bytes"] / n, "dram_write_bytes": a["dram_write_bytes"] / n,
             "grid": a["grid"], "block": a["block"], "regs": a["regs"], "smem_dyn_bytes": a["smem_dyn_bytes"]}
        e["dram_gbs"] = (e["dram_read_bytes"] + e["dram_write_bytes"]) / t / 1e3 if t else None
        e["frac_of_hbm_peak"] = e["dram_gbs"] / peak if t else None
        e["hbm_peak_gbs"] = peak
        e["tensor_pipe_pct"] = sum(a["tensor"]) / len(a["tensor"]) if a["tensor"] else None
        e["issue_active_pct"] = sum(a["issue"]) / len(a["issue"]) if a["issue"] else None
        res.append(e)
    res.sort(key=lambda e: -e["time_us"] * e["launches"])
    json.dump(res, open(out, "w"), indent=1)
    for e in res:
        print(f"{e['kernel'][:60]:60s} n={e['launches']:3d} {e['time_us']:8.1f} us  dram {e['dram_gbs'] or 0:7.0f} GB/s ({100 * (e['frac_of_hbm_peak'] or 0):4.1f}% of {peak:.0f})"
              f"  rd {e['dram_read_bytes'] / 1e6:7.1f} MB wr {e['dram_write_bytes'] / 1e6:7.1f} MB  tensor {e['tensor_pipe_pct']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
