#!/usr/bin/env python
"""One-table digest of the `ncu --set full` captures in gpurun_out/*.ncu-rep -> profiles/ncu_digest.md (runs without a GPU)."""
import csv
import glob
import io
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PICK = [("time", "gpu__time_duration.sum", "us", 1e-3),
        ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", 1),
        ("DRAM read MB", "dram__bytes_read.sum", "MB", None),
        ("DRAM write MB", "dram__bytes_write.sum", "MB", None),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", "%", 1),
        ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "%", 1),
        ("tensor pipe % of SM-active cycles", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%", 1),
        ("tensor pipe % of elapsed", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "%", 1),
        ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active", "%", 1),
        ("regs/thread", "launch__registers_per_thread", "", 1),
        ("grid x block", None, "", None)]


def to_mb(val, unit):
    v = float(val.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    return v * scale


def main():
    rows_out, stalls_out = [], []
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep"))):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=600).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        header, units = rows[0], rows[1]
        unit_of = dict(zip(header, units))
        rec = dict(zip(header, rows[2]))
        name = rec.get("Kernel Name", "?").split("(")[0].replace("void b200::<unnamed>::", "").replace("void b200::", "")
        cells = [f"`{name[:60]}`"]
        for label, key, unit, scale in PICK:
            if key is None:
                cells.append(f"{rec.get('launch__grid_size', '?')} x {rec.get('launch__block_size', '?')}")
                continue
            val = rec.get(key, "")
            if not val:
                cells.append("-")
            elif unit == "MB":
                cells.append(f"{to_mb(val, unit_of.get(key, 'byte')):.1f}")
            elif key == "gpu__time_duration.sum":
                v = float(val.replace(",", ""))
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit_of.get(key, "ns"), 1e-3)
                cells.append(f"{v:.1f}")
            else:
                cells.append(f"{float(val.replace(',', '')):.1f}")
        rows_out.append("| " + " | ".join(cells) + " |")
        stall = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(rec[h].replace(",", "")) for h in header
                 if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and rec.get(h)}
        total = sum(stall.values()) or 1.0
        top = sorted(stall.items(), key=lambda kv: -kv[1])[:4]
        stalls_out.append(f"| `{name[:60]}` | " + ", ".join(f"{k} {100 * v / total:.0f} %" for k, v in top) + " |")
    out = ["# ncu digest: the six `--set full` captures of the round", "",
           "From `gpurun_out/*.ncu-rep` (scratch) via `tools/ncu_digest.py`; full metric dumps per kernel are the `ncu_prof_*.md` files next to this one.",
           "Captured with `--clock-control none` inside `bench/kernel_bench.py` (cold L2: the bench flushes between launches). Never used as a timing source.", "",
           "| kernel | " + " | ".join(f"{l}{(' [' + u + ']') if u and u != '%' else ''}" for l, _, u, _ in PICK) + " |",
           "|" + "---|" * (len(PICK) + 1)] + rows_out + ["", "Top warp-stall reasons (share of sampled issue stalls):", "",
           "| kernel | stalls |", "|---|---|"] + stalls_out + [""]
    path = os.path.join(ROOT, "profiles", "ncu_digest.md")
    open(path, "w").write("\n".join(out))
    print(path)


if __name__ == "__main__":
    main()
