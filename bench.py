#!/usr/bin/env python
"""Benchmark of the HSIMAE pretraining hot path (BASELINE.json metric: pretrain patches/sec, 9x9 HSI, fwd+bwd).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

A "step" is what the reference's training loop does per batch (Model_Pretraining.py:96-106): forward,
zero_grad, backward, AdamW step -- on the HSIMAE-Large config (Model_Pretraining.py:121-131), mask ratio 0.5,
batch 4096 synthetic 9x9x32 patches per GPU (BASELINE.json configs[1]); weak scaling over GPUs.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LARGE = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, embed_dim=256, depth=12, num_heads=16, s_depth=9,
             decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
FLOP_PER_PATCH = 1.893e9          # algorithmic fwd+bwd FLOP / patch, Large, mask 0.5
CUBE = 32 * 9 * 9
METRIC = "pretrain patches/sec (9x9 HSI, fwd+bwd)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sus=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = max(int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def reference_available() -> bool:
    from oracle import fetch_ref
    return fetch_ref.root() is not None


def reference_step_fn(batch: int, device: str):
    """The UNMODIFIED reference (`HSIMAE` of /root/reference/Models.py, or its byte-identical travelling copy in
    oracle/_ref) driven exactly as Model_Pretraining.py:68-106 drives it: Large config, AdamW two groups, forward,
    zero_grad, backward, step, loss.item().  fp32 eager; TF32 off so the arithmetic is the reference's."""
    from oracle import fetch_ref
    R = fetch_ref.import_models()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(42); random.seed(42)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = R.HSIMAE(**LARGE).to(device)
    model.train()
    no_decay = ("bias", "norm")
    groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
              {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=5e-3, weight_decay=5e-2, betas=(0.9, 0.95))
    x = torch.randn(batch, 1, 32, 9, 9, device=device)

    def step():
        loss, _, _ = model(x, mask_ratio=0.5)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()
    return step


def cpu_step_fn(batch: int):
    """one training step of the reference algorithm on the host cores: the unmodified reference when it is available
    (kind "reference"), else the oracle port (kind "port")"""
    if reference_available():
        return reference_step_fn(batch, "cpu"), "reference"
    from oracle import hsimae_oracle as O
    geo = O.Geometry()
    torch.manual_seed(42); random.seed(42)
    sd = O.make_state(geo, seed=42)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in sd.items()}
    no_decay = ("bias", "norm")
    groups = [{"params": [p for n, p in leaves.items() if p.requires_grad and not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
              {"params": [p for n, p in leaves.items() if p.requires_grad and any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=5e-3, weight_decay=5e-2, betas=(0.9, 0.95))
    x = torch.randn(batch, 1, 32, 9, 9)

    def step():
        lt, ll = O.choose_visible_shape(geo.T, geo.L, 0.5)
        out = O.pretrain_forward(leaves, x, geo, torch.rand(batch, geo.T), torch.rand(batch, geo.L), lt, ll)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        return float(out["loss"].detach())
    return step, "port"


def run_cpu(steps: int, warmup: int, batch: int = 64, budget_s: float = 1e9):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = cpu_step_fn(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter(); done = 0
    for _ in range(steps):
        step(); done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=batch * done / dt, ms_per_step=1e3 * dt / done, cores=cores, steps=done, batch=batch, kind=kind)


def run_gpu_eager(batch: int, steps: int = 5, warmup: int = 3):
    """the bar SURVEY 2.1 names: the unmodified reference, PyTorch eager fp32, on the SAME B200 and batch"""
    step = reference_step_fn(batch, "cuda")
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return dict(value=batch / (ms * 1e-3), unit="patches/s", ms_per_step=ms, batch=batch, steps=steps,
                kind="unmodified reference Models.HSIMAE, PyTorch eager fp32 (TF32 off), torch AdamW, loss.item() every step, cuda:0")


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args.steps, max(args.warmup, 1), 64, budget_s=240.0)
    what = "unmodified reference Models.HSIMAE" if r["kind"] == "reference" else "oracle port"
    sample = f"{what}: HSIMAE-Large fwd+bwd+AdamW, batch {r['batch']} synthetic patches x {r['steps']} steps, fp32, {r['cores']} torch threads"
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "patches/s", "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "HSIMAE-Large pretraining step (fwd+bwd+AdamW), mask 0.5, 9x9x32 patches; CPU sample batch 64"},
            "cpu_baseline": {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- dominant-kernel roofline
def time_events(fn, iters: int):
    """average device time of one launch: `iters` back-to-back launches between two events on the launching stream
    (operands + outputs of every timed kernel are several times larger than the 126 MB L2, so nothing is served from cache)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def ncu_traffic(cand: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the candidate kernel, from the committed
    `ncu --set full` capture (profiles/r01e_ncu_full_kernels.json, written by tools/ncu_summary.py from a
    tools/prof_kernels.py run at the shapes timed here), or None"""
    want = {"gemm_tc_dgate_kernel": "d(gate)", "mlp_fused_kernel": "fused gated MLP", "gemm_tc_kernel<Bias> qkv": "qkv projection",
            "wgrad_group_kernel": "wgrad group", "gemm_tc_kernel<LnBwd> dgrad [M,1376]": "dgrad K=1376 + LN bwd",
            "gemm_tc_kernel<LnBwd> dgrad [M,768]": "dgrad K=768 + LN bwd"}
    for name in ("r02i_ncu_full_kernels.json", "r02_ncu_full_kernels.json", "r01e_ncu_full_kernels.json"):
        try:
            rows = json.load(open(os.path.join(ROOT, "profiles", name)))
            for key, label in want.items():
                if cand.startswith(key):
                    hit = [r for r in rows if r["label"].startswith(label)][0]
                    return hit["dram_read_bytes"] + hit["dram_write_bytes"]
        except Exception:
            pass
    return None


def dominant_kernel_roofline(batch: int, pk):
    """Times the per-block GEMM kernels at the bench shapes (M = batch*18 encoder token rows) in isolation, picks the
    class with the largest share of a step, and reports it against its binding roofline.  Algorithmic work per launch
    (DESIGN.md section 3): FLOPs = 2 M N K per contraction; bytes = every operand / result tensor once."""
    from hsimae_b200 import ops
    dev = "cuda"
    M, D, H = batch * 18, 256, 688
    bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
    x, w13, wqkv, w2 = bf(M, D), bf(2 * H, D) * 0.05, bf(3 * D, D) * 0.05, bf(D, H) * 0.05
    w2t = bf(H, D) * 0.05
    g, resid = bf(M, H), torch.randn(M, D, device=dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dab = bf(M, 2 * H)
    gw13 = torch.zeros(684, D, device=dev); gw13b = torch.zeros(684, D, device=dev)
    per_step = 21   # encoder blocks per direction
    x2 = bf(M, D)          # the d(gate) call reads TWO distinct activations (dy and the LayerNorm-2 output)
    w13t_full, wqkv_t, dqkv = bf(D, 2 * H) * 0.05, bf(D, 3 * D) * 0.05, bf(M, 3 * D)
    xf, dxf = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
    stats = torch.stack([xf.mean(1), (xf.var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
    dgam, dbet = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    z = lambda *s_: torch.zeros(*s_, device=dev)
    wg_jobs = [dict(Y=x, X=g, dst0=z(D, 684), cols_valid=684, bias0=z(D)),
               dict(Y=dab, X=x2, dst0=gw13, dst1=gw13b, row_map=1, rows_valid=684, bias0=z(684), bias1=z(684)),
               dict(Y=x, X=x2, dst0=z(D, D), bias0=z(D)),
               dict(Y=dqkv, X=x2, dst0=z(3 * D, D), bias0=z(3 * D))]
    b13, b2 = torch.zeros(2 * H, device=dev), torch.zeros(D, device=dev)
    # name -> (launch, ALGORITHMIC flops per SURVEY 8(d) accounting, algorithmic bytes, launches per step, EXECUTED flops)
    cands = {
        "gemm_tc_dgate_kernel d(a|b) from [M,256]x{[256,1376],[256,688]}": (
            lambda: ops.gemm(x, w2t, ops.EPI_DGATE, A2=x2, B2=w13), 2 * M * D * H, 2 * (2 * M * D + 3 * H * D + M * 2 * H), per_step,
            2 * M * D * 3 * H),   # + the recomputed up-projection (2 M D 2H), which SURVEY 8(d)'s 1.893 GFLOP/patch does not contain
        "mlp_fused_kernel [M,256]x[256,1376] -> gate -> x[688,256] + residual + LayerNorm (training: g kept)": (
            lambda: ops.mlp_fused(x, w13, b13, w2, b2, resid, gamma=gamma, beta=beta), 2 * M * 3 * H * D,
            2 * M * D + 4 * M * D + 2 * 3 * H * D + 2 * M * H + 4 * M * D + 2 * M * D, per_step, 2 * M * 3 * H * D),
        "gemm_tc_kernel<Bias> qkv [M,256]x[256,768]": (lambda: ops.gemm(x, wqkv, ops.EPI_BIAS_BF16), 2 * M * 3 * D * D,
                                                            2 * (M * D + 3 * D * D + M * 3 * D), per_step, 2 * M * 3 * D * D),
        "wgrad_group_kernel dW2 | dW1,dW3 | dWproj | dWq,k,v of one block (reduction over M)": (
            lambda: ops.wgrad_group(wg_jobs), 2 * M * D * (H + 2 * H + D + 3 * D), 2 * M * (D + H + 2 * H + D + D + D + 3 * D + D), per_step,
            2 * M * D * (H + 2 * H + D + 3 * D)),
        "gemm_tc_kernel<LnBwd> dgrad [M,1376]x[1376,256] + LayerNorm backward + residual gradient": (
            lambda: ops.gemm_lnbwd(dab, w13t_full, xf, stats, gamma, dxf, dgamma=dgam, dbeta=dbet, inplace=True),
            2 * M * 2 * H * D, 2 * (M * 2 * H + 2 * H * D) + M * D * (4 + 4 + 4 + 2), per_step, 2 * M * 2 * H * D),
        "gemm_tc_kernel<LnBwd> dgrad [M,768]x[768,256] + LayerNorm backward + residual gradient": (
            lambda: ops.gemm_lnbwd(dqkv, wqkv_t, xf, stats, gamma, dxf, dgamma=dgam, dbeta=dbet, inplace=True),
            2 * M * 3 * D * D, 2 * (M * 3 * D + 3 * D * D) + M * D * (4 + 4 + 4 + 2), per_step, 2 * M * 3 * D * D),
    }
    rows = {}
    for name, (fn, flops, nbytes, count, executed) in cands.items():
        t = time_events(fn, 20)
        rows[name] = dict(t=t, flops=flops, bytes=nbytes, count=count, executed=executed)
    name = max(rows, key=lambda k: rows[k]["t"] * rows[k]["count"])
    r = rows[name]
    t_tensor, t_hbm = r["flops"] / (pk["tf"] * 1e12), r["bytes"] / (pk["hbm"] * 1e9)
    if t_hbm >= t_tensor:
        ach = r["bytes"] / r["t"] / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    else:
        ach = r["flops"] / r["t"] / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"]}
    roof.update(kernel=name, us_per_launch=r["t"] * 1e6, traffic=ncu_traffic(name), peak_source=pk["src"],
                algorithmic_bytes=r["bytes"], algorithmic_flops=r["flops"], executed_flops=r["executed"],
                executed_tflops=r["executed"] / r["t"] / 1e12,
                note="write-only HBM streams on this part peak at ~3.9 TB/s (tools/membw.py, profiles/r01e_membw.txt); "
                     "write-heavy epilogues are bounded by that, not by the 6.55 TB/s copy figure",
                all_kernels={k: {"us": v["t"] * 1e6, "tflops": v["flops"] / v["t"] / 1e12, "executed_tflops": v["executed"] / v["t"] / 1e12,
                                 "gbs": v["bytes"] / v["t"] / 1e9} for k, v in rows.items()})
    return roof



# ----------------------------------------------------------------------------- in-step kernel accounting
GEMM_FLOP_PER_PATCH = 3 * 626.0e6 - 2 * 1.33e6   # algorithmic GEMM FLOP / patch fwd+bwd (no dX for the patch embedding)


def step_kernel_accounting(step_fn, batch: int, pk):
    """One extra (untimed) step under the CUPTI activity profiler: device time of EVERY kernel of the step by name.
    Gives (c) the GEMM-family fraction the north-star target is about -- algorithmic GEMM + weight-gradient FLOPs of a
    step / the summed in-step time of the kernels that execute them / peak -- and (d) the HBM-bound kernels with their
    algorithmic bytes against the measured HBM peak."""
    from torch.profiler import profile, ProfilerActivity
    from hsimae_b200 import _lib
    lib = _lib.load()
    try:
        # plain stream order for this step: with programmatic dependent launch the next kernel's prologue starts (and its
        # activity record opens) while the previous kernel drains, so consecutive durations would overlap
        was = lib.hsimae_set_pdl(0)
        try:
            step_fn()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step_fn()
                torch.cuda.synchronize()
        finally:
            lib.hsimae_set_pdl(was)
        evs = [e for e in prof.events() if getattr(e, "device_time", 0) and "Memcpy" not in e.name and "Memset" not in e.name]
    except Exception as e:   # a reported breakdown, never fatal
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    agg = {}
    for e in evs:
        a = agg.setdefault(e.name, [0, 0.0])
        a[0] += 1; a[1] += e.device_time   # us
    fam = {"gemm+ln_bwd (fused)": ("gemm_tc_kernel<6,",), "gemm": ("gemm_tc", "mlp_fused", "wgrad_"), "attention": ("attn_",), "ln_bwd": ("ln_bwd",), "embed": ("embed_",),
           "loss": ("loss_kernel",), "fill": ("fill_",), "mask": ("mask_kernel",), "pack": ("pack_kernel",), "nccl": ("nccl",)}
    fams = {k: [0, 0.0] for k in fam}; fams["other (optimizer, rand, casts)"] = [0, 0.0]
    for name, (n, us) in agg.items():
        for k, keys in fam.items():
            if any(x in name for x in keys):
                fams[k][0] += n; fams[k][1] += us; break
        else:
            fams["other (optimizer, rand, casts)"][0] += n; fams["other (optimizer, rand, casts)"][1] += us
    total = sum(v[1] for v in fams.values())
    g_pure_us, nfused = fams["gemm"][1], fams["gemm+ln_bwd (fused)"][0]
    g_us = g_pure_us + fams["gemm+ln_bwd (fused)"][1]
    # the fused dgrad + LayerNorm-backward launches (21 encoder blocks x 2 at Large): their GEMM FLOPs, for the fraction without them
    fused_flop = nfused * 2.0 * (batch * 18) * 256 * (2 * 688 + 3 * 256) / 2
    out = {"kernel_time_sum_ms": total / 1e3, "launches": sum(v[0] for v in fams.values()),
           "families": {k: {"launches": v[0], "ms": v[1] / 1e3, "share": v[1] / total} for k, v in fams.items() if v[0]},
           "note": "one untimed step in plain stream order (programmatic dependent launch off) under the CUPTI activity profiler",
           "gemm_family_frac": {"algorithmic_tflop_per_step": batch * GEMM_FLOP_PER_PATCH / 1e12, "in_step_ms": g_us / 1e3,
                                "achieved_tflops": batch * GEMM_FLOP_PER_PATCH / (g_us * 1e-6) / 1e12, "peak_tflops": pk["tf_sus"],
                                "frac": batch * GEMM_FLOP_PER_PATCH / (g_us * 1e-6) / 1e12 / pk["tf_sus"], "peak": "sustained, " + pk["src"],
                                "frac_of_burst": batch * GEMM_FLOP_PER_PATCH / (g_us * 1e-6) / 1e12 / pk["tf"],
                                "frac_excl_fused_ln": (batch * GEMM_FLOP_PER_PATCH - fused_flop) / (g_pure_us * 1e-6) / 1e12 / pk["tf_sus"],
                                "note": "GEMM + weight-gradient kernels only (tcgen05): algorithmic FLOPs, recomputation not counted; the "
                                        "dgrad kernels that also do the LayerNorm backward (HBM-bound work of the former ln_bwd family) are "
                                        "inside `frac` and left out of `frac_excl_fused_ln`"}}
    # HBM-bound kernels: algorithmic bytes per launch (every operand / result once), Large, mask 0.5
    B, K, P, D, Dd, PK = batch, 18, 36, 256, 64, 72
    M, Md = B * K, B * P
    byts = {"embed_fwd": B * K * PK * 4 + M * D * 4 + 2 * M * D * 2, "embed_bwd": B * K * PK * 4 + 2 * M * D * 4,
            "loss_kernel": B * CUBE * 4 + Md * 80 * 4 + Md * 80 * 2 + 2 * B * CUBE * 4, "fill_fwd": M * Dd * 4 + Md * Dd * 4 + Md * Dd * 2,
            "fill_bwd": Md * Dd * 4 + M * Dd * 2, "mask_kernel": B * (13 * 4 + (18 + 36) * 8 + 36 * 4 + 54 * 4),
            "ln_bwd_vec_kernel<8>": M * D * (2 + 4 + 4 + 4 + 2), "ln_bwd_vec_kernel<2>": Md * Dd * (2 + 4 + 4 + 4 + 2)}
    mem = {}
    for name, (n, us) in agg.items():
        for key, nb in byts.items():
            if key in name:
                t = us / n * 1e-6
                mem[key] = {"launches": n, "us": us / n, "algorithmic_bytes": nb, "gbs": nb / t / 1e9, "frac_of_hbm_peak": nb / t / 1e9 / pk["hbm"]}
    out["membound_kernels"] = mem
    out["top_kernels"] = [{"name": k[:90], "launches": v[0], "ms": v[1] / 1e3} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]]
    return out

# ----------------------------------------------------------------------------- secondary workloads
def feed_main(args):
    """`--workload feed`: on-device pretraining data feed (hsimae_gather_patches) on a Salinas-shaped scene, batch 4096,
    beside the oracle restatement of the reference's per-sample host loop (Model_Pretraining.py:40-51) as cpu_baseline."""
    import numpy as np
    from hsimae_b200.feed import PatchFeed
    from oracle import feed_oracle as FO
    rng = np.random.default_rng(0)
    scene = rng.standard_normal((512, 217, 32)).astype(np.float32)
    cut = np.array([(0, h, w, 0, 1, 0) for h in range(0, 504) for w in range(0, 209)], dtype=np.int16)
    feed = PatchFeed([[scene], cut], train=True)
    B = args.batch
    idx = torch.randint(0, len(cut), (B,))
    flips = torch.randint(0, 2, (B, 2), dtype=torch.uint8)
    for _ in range(max(args.warmup, 3)):
        feed.batch(idx, flips)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        x = feed.batch(idx, flips)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    nbytes = 2 * B * CUBE * 4
    t0 = time.perf_counter()
    ref = FO.get_batch([scene], cut, idx[:512].numpy(), flips[:512].numpy())
    dt = time.perf_counter() - t0
    assert np.array_equal(x[:512].cpu().numpy(), ref), "device feed differs from the oracle"
    pk = peaks()
    print(json.dumps({"metric": "pretraining data feed patches/sec (9x9x32 windows from an HBM-resident scene)", "value": B / (ms * 1e-3),
                      "unit": "patches/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"PatchFeed.batch, Salinas-shaped scene 512x217x32, batch {B}, flips on; includes the index/flip H2D copies"},
                      "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": nbytes / ms / 1e6 / pk["hbm"], "traffic": None, "peak_source": pk["src"]},
                      "cpu_baseline": {"value": 512 / dt, "unit": "patches/s", "cores": 1, "kind": "port",
                                       "sample": "512 samples through the numpy restatement of HSIdataset4PT.__getitem__ + collation"}}), flush=True)


def gwpca_main(args):
    """`--workload gwpca`: group-wise PCA of a Salinas-sized raw scene (512 x 217 x 204 float64) on the device, beside the
    numpy oracle (applyGWPCA restated, Utils/GroupWisePCA.py:20-34) on the host cores as cpu_baseline."""
    import numpy as np
    from hsimae_b200.gwpca import applyGWPCA, _auto_sign
    from oracle import gwpca_oracle as G
    rng = np.random.default_rng(0)
    H, W, Cb = 512, 217, 204
    X = rng.normal(size=(H, W, 8)) @ rng.normal(size=(8, Cb)) * 300 + 40 * rng.normal(size=(H, W, Cb)) + 5000
    xd = torch.from_numpy(X).cuda()
    for _ in range(max(args.warmup, 3)):
        out = applyGWPCA(xd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = applyGWPCA(xd)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = applyGWPCA(X)
    torch.cuda.synchronize()
    t_h2d = (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    ref = G.apply_gwpca(X, sign=_auto_sign())
    t_cpu = time.perf_counter() - t0
    err = float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 1e-6, f"device GWPCA differs from the oracle: {err}"
    nbytes = 3 * X.nbytes + out.numel() * 8
    pk = peaks()
    print(json.dumps({"metric": "group-wise PCA scenes/sec (512x217x204 float64 -> 32 whitened components)", "value": 1.0 / t_dev,
                      "unit": "scenes/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev * 1e3,
                      "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "applyGWPCA(nc=32, group=4, whiten=True), scene resident in HBM; host part = 2 D2H reads + 4 eigh(51x51)"},
                      "e2e": {"value": 1.0 / t_h2d, "unit": "scenes/s", "h2d_bytes_per_step": X.nbytes, "d2h_bytes_per_step": 4 * 64 * 64 * 8 + 16,
                              "ms_per_step": t_h2d * 1e3},
                      "roofline": {"bound": "hbm", "achieved": nbytes / t_dev / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": nbytes / t_dev / 1e9 / pk["hbm"], "traffic": None, "peak_source": pk["src"],
                                   "note": "whole call incl. the host eigen-solves and two synchronising reads, not one kernel"},
                      "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                                       "sample": "one scene through oracle/gwpca_oracle.py (numpy / LAPACK)"},
                      "max_rel_err_vs_oracle": err}), flush=True)



# ----------------------------------------------------------------------------- BASELINE.json configs[3] / configs[4]
DUAL = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=256, depth=12, num_heads=16, s_depth=9,
            drop_path=0.2, decoder_embed_dim=64, decoder_depth=8, decoder_num_heads=8, norm_pix_loss=True, trunc_init=True)
VIT = dict(img_size=9, patch_size=3, in_chans=1, bands=32, b_patch_size=8, num_class=17, embed_dim=256, depth=12, num_heads=16, s_depth=9,
           trunc_init=True)
FT_FLOP_PER_STEP = 3 * (32 * 1200.3e6 + 103 * 298.3e6)   # unmasked encoder on 32 labelled + masked branch on 103, fwd+bwd
SCENE_FLOP = 512 * 217 * 1200.3e6                        # encoder-only forward over every pixel-centred window


def _events(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, r


def finetune_main(args):
    """`--workload finetune` (BASELINE.json configs[3]): the dual-branch fine-tuning step of Model_Finetuning.py:147-166 --
    DualViT-Large, 32 labelled + 71 unlabelled Salinas-shaped tiles, mask 0.8, lambda 10, drop_path 0.2, CE(ignore_index=0),
    AdamW -- device-resident and end to end (x, x_u, y from pinned host memory and the logits read back every step, as the
    driver does at :150-156); the unmodified reference on the host cores beside it."""
    import Models
    from hsimae_b200 import _lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(0); random.seed(0)
    model = Models.DualViT(**DUAL).to(dev).train()
    if args.fused_optimizer:
        from hsimae_b200.optim import FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=5e-2)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-2)
    crit = torch.nn.CrossEntropyLoss(ignore_index=0)
    nl, nu = args.labelled, args.unlabelled
    hx, hxu = torch.randn(nl, 1, 32, 9, 9).pin_memory(), torch.randn(nu, 1, 32, 9, 9).pin_memory()
    hy = torch.randint(1, 17, (nl,)).pin_memory()
    x, xu, y = hx.to(dev), hxu.to(dev), hy.to(dev)
    lib = _lib.load()
    stepper = None
    if not args.no_graph:
        try:
            from hsimae_b200.graph import GraphedFinetuneStep
            stepper = GraphedFinetuneStep(model, opt, crit, lamda=10.0, mask_ratio=0.8)
        except Exception as e:   # the eager step is always available
            print(f"# CUDA-graph step unavailable: {type(e).__name__}: {e}", file=sys.stderr)

    def step_eager(a, b, c):
        loss_rec, _, _, logits = model(a, b, mask_ratio=0.8)
        loss = 10 * loss_rec + crit(logits, c)
        opt.zero_grad(); loss.backward(); opt.step()
        return loss, logits

    step = (lambda a, b, c: stepper(a, b, c)) if stepper is not None else step_eager
    l0 = lib.hsimae_launch_count()
    ms, _ = _events(lambda: step(x, xu, y), args.steps, max(args.warmup, 3))
    launches = (lib.hsimae_launch_count() - l0) / (args.steps + max(args.warmup, 3))

    def e2e_step():
        a, b, c = hx.to(dev, non_blocking=True), hxu.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)
        loss, logits = step(a, b, c)
        return logits.detach().cpu()          # Model_Finetuning.py:156
    e2e_ms, _ = _events(e2e_step, args.steps, 2)
    pk = peaks()
    tf = FT_FLOP_PER_STEP / (ms * 1e-3) / 1e12
    line = {"metric": "dual-branch fine-tuning patches/sec (DualViT-Large, fwd+bwd+AdamW)", "value": (nl + nu) / (ms * 1e-3), "unit": "patches/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"Model_Finetuning.py:147-166 step, {nl} labelled + {nu} unlabelled 9x9x32 tiles, 17 classes, mask 0.8, "
                                   "lambda 10, drop_path 0.2 (BASELINE.json configs[3])",
                       "execution": "one CUDA graph per visible shape (hsimae_b200.graph)" if stepper is not None else "eager launches",
                       "optimizer": "hsimae_b200.optim.FusedAdamW (opt-in)" if args.fused_optimizer else "torch.optim.AdamW (driver-owned, unchanged)"},
            "e2e": {"value": (nl + nu) / (e2e_ms * 1e-3), "unit": "patches/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": (nl + nu) * CUBE * 4 + nl * 8, "d2h_bytes_per_step": nl * 17 * 4},
            "gpu_launches_per_step": launches,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tf / pk["tf_sus"], "traffic": None,
                         "peak_source": "sustained, " + pk["src"],
                         "note": "whole step (~1000 kernels on 103 samples): launch / latency bound, not a kernel roofline"}}
    if not args.no_cpu_baseline and reference_available():
        from oracle import fetch_ref
        R = fetch_ref.import_models()
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0); random.seed(0)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            ref = R.DualViT(**DUAL).train()
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=5e-2)
        cx, cxu, cy = hx.clone(), hxu.clone(), hy.clone()

        def cpu_step():
            loss_rec, _, _, logits = ref(cx, cxu, mask_ratio=0.8)
            loss = 10 * loss_rec + crit(logits, cy)
            ropt.zero_grad(); loss.backward(); ropt.step()
        cpu_step()
        t0 = time.perf_counter(); n = 0
        while n < 5 and time.perf_counter() - t0 < 20:
            cpu_step(); n += 1
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": (nl + nu) / dt, "unit": "patches/s", "cores": os.cpu_count(), "kind": "reference",
                                "sample": f"unmodified reference DualViT, same batch, {n} steps, fp32 ({dt * 1e3:.0f} ms/step)"}
    print(json.dumps(line), flush=True)


def scene_main(args):
    """`--workload scene` (BASELINE.json configs[4]): dense per-pixel classification of a Salinas-sized synthetic scene
    (512 x 217 x 32) with HSIViT-Large, 9x9 windows gathered on the device (Model_Finetuning.py:243-301 semantics)."""
    import Models
    from hsimae_b200.scene import classify_scene
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    vit = Models.HSIViT(**VIT).to(dev).eval()
    hscene = torch.randn(512, 217, 32).pin_memory()
    scene = hscene.to(dev)
    batch = args.scene_batch
    ms, _ = _events(lambda: classify_scene(vit, scene, batch=batch), max(args.steps // 4, 3), 2)

    def e2e():
        out = classify_scene(vit, hscene.to(dev, non_blocking=True), batch=batch)
        return (out[:, 1:].argmax(1) + 1).to(torch.int16).cpu()     # Model_Finetuning.py:277-280
    e2e_ms, labels = _events(e2e, max(args.steps // 4, 3), 1)
    pk = peaks()
    npx = 512 * 217
    tf = SCENE_FLOP / (ms * 1e-3) / 1e12
    line = {"metric": "dense scene classification pixels/sec (HSIViT-Large, 512x217x32, 9x9 windows)", "value": npx / (ms * 1e-3), "unit": "pixels/s",
            "n_gpus": 1, "steps": max(args.steps // 4, 3), "warmup": 2, "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"whole-scene inference, on-device sliding windows, batches of {batch} windows (BASELINE.json configs[4])"},
            "e2e": {"value": npx / (e2e_ms * 1e-3), "unit": "pixels/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": hscene.numel() * 4,
                    "d2h_bytes_per_step": npx * 2},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tf / pk["tf_sus"], "traffic": None,
                         "peak_source": "sustained, " + pk["src"], "note": "encoder-only forward, 1200.3 MFLOP per window (SURVEY 8d)"}}
    if not args.no_cpu_baseline and reference_available():
        from oracle import fetch_ref
        R = fetch_ref.import_models()
        torch.set_num_threads(os.cpu_count() or 1)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            ref = R.HSIViT(**VIT).eval()
        xb = torch.randn(256, 1, 32, 9, 9)
        with torch.no_grad():
            ref(xb)
            t0 = time.perf_counter(); n = 0
            while n < 8 and time.perf_counter() - t0 < 20:
                ref(xb); n += 1
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": 256 / dt, "unit": "pixels/s", "cores": os.cpu_count(), "kind": "reference",
                                "sample": f"unmodified reference HSIViT on host cubes, batch 256 (Model_Finetuning.py:265) x {n}, fp32"}
    print(json.dumps(line), flush=True)


def dp_parity_check(model, rank: int, world: int, dev):
    """N-rank averaged gradients == 1-rank gradients on the concatenated batch (one shot, before timing): every rank runs
    the global parity batch alone, then its shard with the NCCL exchange attached; same noise, same visible shape."""
    import torch.distributed as dist
    import hsimae_b200.modules as mod
    from hsimae_b200 import dp
    per = 16
    g = torch.Generator(device="cpu").manual_seed(11)
    x_all = torch.randn(world * per, 1, 32, 9, 9, generator=g).to(dev)
    nt_all, nl_all = torch.rand(world * per, 4, generator=g).to(dev), torch.rand(world * per, 9, generator=g).to(dev)

    def run(x, nt, nl):
        feed = [nt.contiguous(), nl.contiguous()]
        orig_s, orig_r = mod.choose_visible_shape, torch.rand
        mod.choose_visible_shape = lambda T, L, r: (3, 6); torch.rand = lambda *a, **k: feed.pop(0)
        try:
            model.zero_grad(); loss, _, _ = model(x, mask_ratio=0.5)
        finally:
            mod.choose_visible_shape, torch.rand = orig_s, orig_r
        loss.backward()
        return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    full = run(x_all, nt_all, nl_all)
    sync = dp.attach(model)
    sl = slice(rank * per, (rank + 1) * per)
    part = run(x_all[sl], nt_all[sl], nl_all[sl])
    worst = max(float((part[k] - full[k]).norm() / (full[k].norm() + 1e-12)) for k in full if not k.endswith("attn.k.bias"))
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    model.zero_grad(set_to_none=True)
    if t.item() > 5e-3:
        raise SystemExit(f"data-parallel gradients differ from the single-replica gradients: max rel {t.item():.3e}")
    return float(t.item()), sync

# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="patches per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-PyTorch-eager-on-this-GPU baseline")
    ap.add_argument("--profile", action="store_true", help="bracket the timed steps with cudaProfilerStart/Stop (for ncu)")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune", "scene", "feed", "gwpca"],
                    help="pretrain = the BASELINE.json metric (default); finetune / scene = BASELINE.json configs[3] / configs[4]; "
                         "feed / gwpca = the SURVEY 8(f) preprocessing paths; all but pretrain run on 1 GPU")
    ap.add_argument("--labelled", type=int, default=32)
    ap.add_argument("--unlabelled", type=int, default=71)
    ap.add_argument("--scene-batch", type=int, default=16384)
    ap.add_argument("--no-graph", action="store_true", help="finetune: eager launches instead of the captured CUDA graph")
    ap.add_argument("--fused-optimizer", action="store_true",
                    help="opt-in hsimae_b200.optim.FusedAdamW instead of the driver's torch.optim.AdamW (SURVEY 8f-3; not the default metric)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    if args.workload != "pretrain":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl ours) needs a CUDA device; the product path has no CPU fallback")
        return {"feed": feed_main, "gwpca": gwpca_main, "finetune": finetune_main, "scene": scene_main}[args.workload](args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import Models
    from hsimae_b200 import _lib, dp

    torch.manual_seed(42); random.seed(42)
    model = Models.HSIMAE(**LARGE).to(dev)
    model.train()
    dp_parity = None
    if world > 1:
        dp.broadcast_parameters(model)
        dp_parity, _ = dp_parity_check(model, rank, world, dev)   # attaches the gradient exchange
    no_decay = ["bias", "norm"]
    groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 5e-2},
              {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    if args.fused_optimizer:
        from hsimae_b200.optim import FusedAdamW
        opt = FusedAdamW(groups, lr=5e-3, weight_decay=5e-2, betas=(0.9, 0.95))
    else:
        opt = torch.optim.AdamW(groups, lr=5e-3, weight_decay=5e-2, betas=(0.9, 0.95))
    torch.manual_seed(1000 + rank); random.seed(7)   # python RNG identical on all ranks (same visible shape), data differs
    B = args.batch
    pool = [torch.randn(B, 1, 32, 9, 9, device=dev) for _ in range(4)]
    host_pool = [torch.randn(B, 1, 32, 9, 9).pin_memory() for _ in range(4)]
    lib = _lib.load()

    def step(x):
        loss, _, _ = model(x, mask_ratio=0.5)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident inputs ("value")
    for i in range(args.warmup):
        step(pool[i % 4])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.hsimae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        loss = step(pool[i % 4])
    e1.record()
    barrier()
    if args.profile:
        torch.cuda.profiler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = lib.hsimae_launch_count() - launches0
    last_loss = float(loss.item())

    # ---- end to end: pinned host batches, H2D every step (prefetched on a copy stream), loss read back every step
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty(B, 1, 32, 9, 9, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])
            bufs[j].copy_(host_pool[i % 4], non_blocking=True)
            ready[j].record(copy_stream)

    def e2e_loop(n):
        for c in consumed:
            c.record()
        prefetch(0)
        tot = 0.0
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            l = step(bufs[i % 2])
            consumed[i % 2].record()
            tot += l.item()              # device -> host read of the step's result, every step
        return tot

    if args.no_e2e:
        e2e_ms = wall_ms = float("nan")
    else:
        e2e_loop(2)
        barrier()
        e0.record()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0)) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    pk = peaks()
    line = None
    if rank == 0:
        value = world * B / (ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "HSIMAE-Large pretraining step (fwd+bwd+AdamW as Model_Pretraining.py:96-106), mask 0.5, "
                                       f"batch {B} synthetic 9x9x32 patches per GPU (BASELINE.json configs[1])",
                           "global_batch": world * B, "parallelism": f"dp{world}",
                           "l2": "4 rotating input batches; per-step activation working set (~15 GB) >> 126 MB L2",
                           "optimizer": "hsimae_b200.optim.FusedAdamW (opt-in)" if args.fused_optimizer else "torch.optim.AdamW (driver-owned, unchanged)"},
                "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": B * CUBE * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms, "wall_ms_per_step": wall_ms / args.steps},
                "dp_parity_max_rel": dp_parity, "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps, "loss": last_loss,
                "clocks": clocks,
                "step_tensor_frac": {"achieved_tflops": B * FLOP_PER_PATCH / (ms * 1e-3) / 1e12, "peak_tflops": pk["tf_sus"],
                                     "frac": B * FLOP_PER_PATCH / (ms * 1e-3) / 1e12 / pk["tf_sus"], "peak": "sustained, " + pk["src"]}}
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.no_roofline:
        line["kernel_accounting"] = step_kernel_accounting(lambda: step(pool[0]), B, pk)
        del pool, bufs
        opt.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        line["roofline"] = dominant_kernel_roofline(B, pk)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu(steps=10, warmup=2, batch=64, budget_s=20.0)
        what = "unmodified reference Models.HSIMAE" if r["kind"] == "reference" else "oracle port"
        line["cpu_baseline"] = {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": f"same step on the host: batch {r['batch']} x {r['steps']} steps, fp32 {what}, "
                                          f"{r['cores']} torch threads ({r['ms_per_step']:.0f} ms/step)"}
        if reference_available() and not args.no_gpu_eager:
            try:
                line["gpu_eager_baseline"] = run_gpu_eager(B)
            except Exception as e:   # e.g. out of memory on a smaller part: a reported baseline, never fatal
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
