#!/usr/bin/env python
"""Turn gpurun_out/ artefacts (ncu reports, launch lists, kernel_bench json) into the tracked summaries under profiles/.
Runs on the CPU box: `ncu -i` needs no GPU."""
import csv
import glob
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def ncu_raw(rep):
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=300).stdout
    except Exception as exc:
        return None, str(exc)
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return None, txt[:500]
    return rows, None


def summarize_rep(rep):
    rows, err = ncu_raw(rep)
    name = os.path.basename(rep).replace(".ncu-rep", "")
    path = os.path.join(OUT, f"ncu_{name}.md")
    with open(path, "w") as f:
        f.write(f"# ncu --set full: {name}\n\nsource report: `gpurun_out/{os.path.basename(rep)}` (scratch, not tracked); "
                f"captured with `--clock-control none --import-source on`.\n\n")
        if rows is None:
            f.write(f"could not read report: {err}\n")
            return path
        header, units = rows[0], rows[1]
        for data in rows[2:]:
            rec = dict(zip(header, data))
            f.write(f"## {rec.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---|---|\n")
            for h, u in zip(header, units):
                if any(h.startswith(k) or k in h for k in KEYS) or "tensor" in h or h.startswith("dram__") or "stall" in h.lower():
                    f.write(f"| {h} | {rec.get(h, '')} | {u} |\n")
            f.write("\n")
    return path


def summarize_launches(csv_path, title):
    if not os.path.isfile(csv_path):
        return None
    lines = [l for l in open(csv_path, errors="ignore") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows:
        if "gpu__time_duration" not in r.get("Metric Name", ""):
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
        k = r["Kernel Name"]
        agg[k][0] += 1
        agg[k][1] += ns
        total += ns
    if not agg:
        return None
    path = os.path.join(OUT, os.path.basename(csv_path).replace(".csv", ".md"))
    ours = ("b200::", "bucket_allreduce", "gemm_bf16_kernel", "multi_sgd", "multi_sqnorm", "clip_coef", "mse_fwd_bwd", "xent_",
            "layernorm_", "small_linear", "normalize_cl", "peer_")
    with open(path, "w") as f:
        f.write(f"# {title}\n\nsource: `gpurun_out/{os.path.basename(csv_path)}` (ncu --metrics gpu__time_duration.sum, serialised, "
                f"cold caches: compare shares, not absolutes)\n\n")
        n = sum(v[0] for v in agg.values())
        mine = sum(v[1] for k, v in agg.items() if any(o in k for o in ours))
        nm = sum(v[0] for k, v in agg.items() if any(o in k for o in ours))
        f.write(f"{n} launches, {total / 1e6:.3f} ms summed kernel time in the profiled range; b200ddp kernels: {nm} launches, "
                f"{mine / 1e6:.3f} ms ({100 * mine / total:.1f}%)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            tag = " **(ours)**" if any(o in k for o in ours) else ""
            f.write(f"| `{k[:110]}`{tag} | {c} | {ns / 1e3:.1f} | {100 * ns / total:.1f}% |\n")
    return path


def main():
    os.makedirs(OUT, exist_ok=True)
    made = []
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep"))):
        made.append(summarize_rep(rep))
    for name, title in (("launches.csv", "Launch list: one ResNet-50 bf16 step (batch 32, eager)"),
                        ("launches_graph.csv", "Launch list: one ResNet-50 bf16 step (batch 32, CUDA-graph replay)")):
        p = summarize_launches(os.path.join(SRC, name), title)
        if p:
            made.append(p)
    kj = os.path.join(SRC, "kernels.json")
    if os.path.isfile(kj):
        d = json.load(open(kj))
        path = os.path.join(OUT, "kernel_roofline.md")
        with open(path, "w") as f:
            pk = d["peaks"]
            f.write(f"# Per-kernel timings vs the measured roofline\n\npeaks ({pk['source']}): HBM copy {pk['hbm_gbs']} GB/s, cuBLAS bf16 "
                    f"{pk['bf16_tflops_burst']} TFLOP/s burst / {pk['bf16_tflops_sustained']} sustained.  CUDA events, L2 flushed between "
                    f"launches, median of 20 (`bench/kernel_bench.py`).\n\n| kernel | ms | achieved | % of measured peak | library ms | ours/lib time |\n|---|---|---|---|---|---|\n")
            for r in d["rows"]:
                rate = f"{r['tflops']:.1f} TFLOP/s" if "tflops" in r else f"{r.get('gbs', 0):.0f} GB/s"
                lib = f"{r['lib_ms']:.3f}" if r.get("lib_ms") else "-"
                ratio = f"{r['ms'] / r['lib_ms']:.2f}" if r.get("lib_ms") else "-"
                f.write(f"| {r['kernel']} | {r['ms']:.3f} | {rate} | {100 * r.get('frac_of_peak', 0):.1f}% | {lib} | {ratio} |\n")
        made.append(path)
    print("\n".join(str(m) for m in made if m))


if __name__ == "__main__":
    main()
