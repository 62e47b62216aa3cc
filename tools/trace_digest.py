#!/usr/bin/env python
"""Per-rank digest of the CUPTI traces written by `bench.py --trace_dir DIR`: where does one training step spend its time
on each rank, and how much of the gradient communication is exposed?

  python tools/trace_digest.py DIR [--out profiles/ddp_timeline_n2.md] [--label "2 GPUs, default"]

For the LAST complete step of every rank (steps are delimited by the input kernel `normalize_cl_kernel`, the first launch of a
step) it reports: the span of the step, every communication kernel (offset from step start, duration, stream), the end of
backward (last kernel before the optimizer), the gap between the last compute kernel of backward and the clip / SGD kernels
(= exposed communication tail), and idle gaps > 3 us on the compute stream with their neighbours."""
import argparse
import glob
import json
import os
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    name = name.replace("b200::<unnamed>::", "").replace("b200::(anonymous namespace)::", "")
    return name[:70]


def digest(path):
    data = json.load(open(path))
    ev = [e for e in data.get("traceEvents", []) if e.get("cat") == "kernel" and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "normalize_cl_kernel" in e["name"]]
    if len(starts) < 2:
        return None
    a, b = starts[-2], starts[-1]
    step = ev[a:b]
    t0 = step[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in step)
    streams = {}
    for e in step:
        streams.setdefault(e["args"].get("stream", 0), []).append(e)
    main = max(streams.values(), key=len)
    main_id = main[0]["args"].get("stream", 0)
    comm = [e for e in step if e["args"].get("stream", 0) != main_id]
    is_comm = lambda e: ("bucket_allreduce" in e["name"] or "peer_broadcast" in e["name"] or "nccl" in e["name"].lower())   # noqa: E731
    comm = [e for e in step if is_comm(e)]
    opt_i = next((i for i, e in enumerate(main) if "multi_sqnorm" in e["name"] or "clip_coef" in e["name"] or "multi_sgd" in e["name"]), None)
    out = {"file": os.path.basename(path), "step_us": end - t0, "kernels": len(step), "main_stream": main_id,
           "busy_main_us": sum(e["dur"] for e in main), "comm": [], "gaps": []}
    for e in comm:
        out["comm"].append({"name": short(e["name"]), "start_us": round(e["ts"] - t0, 1), "dur_us": round(e["dur"], 1), "stream": e["args"].get("stream", 0),
                            "grid": e["args"].get("grid")})
    if opt_i is not None and opt_i > 0:
        last_bwd = main[opt_i - 1]
        opt = main[opt_i]
        out["backward_end_us"] = round(last_bwd["ts"] + last_bwd["dur"] - t0, 1)
        out["optimizer_start_us"] = round(opt["ts"] - t0, 1)
        out["exposed_tail_us"] = round(opt["ts"] - (last_bwd["ts"] + last_bwd["dur"]), 1)
        out["last_backward_kernel"] = short(last_bwd["name"])
        if comm:
            out["last_comm_end_us"] = round(max(e["ts"] + e["dur"] for e in comm) - t0, 1)
    prev = None
    for e in main:
        if prev is not None:
            gap = e["ts"] - (prev["ts"] + prev["dur"])
            if gap > 3.0:
                out["gaps"].append({"at_us": round(prev["ts"] + prev["dur"] - t0, 1), "gap_us": round(gap, 1), "after": short(prev["name"]), "before": short(e["name"])})
        prev = e
    out["gap_total_us"] = round(sum(g["gap_us"] for g in out["gaps"]), 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--out", default=None)
    ap.add_argument("--label", default="")
    args = ap.parse_args()
    rows = []
    for path in sorted(glob.glob(os.path.join(args.dir, "rank*.json"))):
        d = digest(path)
        if d:
            rows.append(d)
    lines = [f"## DDP step timeline: {args.label or args.dir}", "",
             "CUPTI kernel records of the last complete step per rank (`bench.py --trace_dir`, `tools/trace_digest.py`); a profiler run is never a timing source - "
             "it shows ORDER and OVERLAP.", ""]
    for d in rows:
        lines += [f"### {d['file']}: step {d['step_us']:.0f} us, {d['kernels']} kernels, compute stream busy {d['busy_main_us']:.0f} us, "
                  f"idle gaps > 3 us on it: {d['gap_total_us']:.0f} us", ""]
        if "backward_end_us" in d:
            lines += [f"* backward ends at {d['backward_end_us']} us (`{d['last_backward_kernel']}`), optimizer starts at {d['optimizer_start_us']} us: "
                      f"**exposed tail {d['exposed_tail_us']} us**; last communication kernel ends at {d.get('last_comm_end_us', '-')} us", ""]
        lines += ["| communication kernel | start [us] | duration [us] | grid |", "|---|---|---|---|"]
        for c in d["comm"]:
            lines.append(f"| `{c['name']}` | {c['start_us']} | {c['dur_us']} | {c['grid']} |")
        lines += ["", "| idle gap on the compute stream | at [us] | after | before |", "|---|---|---|---|"]
        for g in sorted(d["gaps"], key=lambda g: -g["gap_us"])[:12]:
            lines.append(f"| {g['gap_us']} us | {g['at_us']} | `{g['after']}` | `{g['before']}` |")
        lines.append("")
    text = "\n".join(lines)
    if args.out:
        with open(args.out, "a") as f:
            f.write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
