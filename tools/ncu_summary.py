"""Summarise an `ncu --set full` report of tools/prof_kernels.py into JSON (one entry per captured launch).

usage: python tools/ncu_summary.py gpurun_out/prof_kernels.ncu-rep profiles/r01c_ncu_full_kernels.json
The labels are the launch order of tools/prof_kernels.py (torch's own fill kernels are skipped)."""
import csv, io, json, subprocess, sys

LABELS = ["qkv projection (streaming pair kernel, bias)", "gated up-projection, g only (A-resident, SwiGLU)", "down-projection + residual + LayerNorm (TMA-staged residual)",
          "d(gate) with recomputed a|b (dual GEMM)", "dgrad K=1376 (pair)", "wgrad dW13 (pair)", "wgrad dW2",
          "attention fwd fusion (len 18)", "attention bwd fusion (len 18)", "attention fwd spatial", "attention fwd spectral",
          "fused gated MLP, training (g kept)", "fused gated MLP, inference",
          "dgrad K=1376 + LN bwd (kEpiLnBwd)", "dgrad K=768 + LN bwd (kEpiLnBwd)", "wgrad group (dW2 | dW13 | dWproj | dWqkv)"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=1.0):
        if name not in col:
            return None
        try:
            v = float(r[col[name]].replace(",", ""))
        except ValueError:
            return None
        u = units[col[name]]
        if u == "Mbyte": v *= 1e6
        elif u == "Kbyte": v *= 1e3
        elif u == "Gbyte": v *= 1e9
        elif u == "ms": v *= 1e3
        elif u == "ns": v *= 1e-3
        return v * scale

    res, li = [], 0
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if "hsimae" not in name:
            continue
        sec = lambda n: (val(r, n) or 0.0) * 32.0
        e = {"label": LABELS[li] if li < len(LABELS) else "", "kernel": name.split("(")[0].replace("void ", ""),
             "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "cluster": r[col["launch__cluster_size"]] if "launch__cluster_size" in col else None,
             "time_us": val(r, "gpu__time_duration.sum"),
             "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
             "l2_sm_read_bytes": sec("lts__t_sectors_srcunit_tex_op_read.sum"), "l2_sm_write_bytes": sec("lts__t_sectors_srcunit_tex_op_write.sum"),
             "l2_total_bytes": sec("lts__t_sectors.sum"),
             "tensor_pipe_pct": None,
             "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "regs": val(r, "launch__registers_per_thread"), "smem_dyn_bytes": val(r, "launch__shared_mem_per_block_dynamic"),
             "warp_inst": val(r, "smsp__inst_executed.sum")}
        for k in col:
            if k.startswith("sm__inst_executed_pipe_tensor") and k.endswith("pct_of_peak_sustained_active") and "hmma" in k:
                e["tensor_inst_pct"] = val(r, k)
            if k == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" or k.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                e["tensor_pipe_pct"] = val(r, k)
        if e["time_us"]:
            e["dram_gbs"] = ((e["dram_read_bytes"] or 0) + (e["dram_write_bytes"] or 0)) / e["time_us"] / 1e3
            e["l2_tbs"] = e["l2_total_bytes"] / e["time_us"] / 1e6
        res.append(e)
        li += 1
    json.dump(res, open(out, "w"), indent=1)
    for e in res:
        print("%-52s %7.1f us  dram R %6.1f W %6.1f MB  L2<->SM R %6.1f W %6.1f MB  tensor %s%%" % (
            e["label"][:52], e["time_us"], (e["dram_read_bytes"] or 0) / 1e6, (e["dram_write_bytes"] or 0) / 1e6,
            e["l2_sm_read_bytes"] / 1e6, e["l2_sm_write_bytes"] / 1e6, e.get("tensor_pipe_pct")))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
