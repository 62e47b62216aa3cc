#!/usr/bin/env python
"""Launch list of the last complete captured step of a `bench.py --trace_dir` CUPTI trace, as markdown: one row per kernel
(short name), launches, total / mean microseconds, share, ours vs library.  A trace is not a timing source: shares only."""
import collections
import json
import re
import sys


def short(name):
    n = name.replace("void ", "")
    n = re.sub(r"b200::\(anonymous namespace\)::|b200::<unnamed>::", "b200::", n)
    n = re.sub(r"\(.*$", "", n)
    n = re.sub(r"<.*$", lambda m: m.group(0) if "conv_tap_gemm" in n or "bucket_allreduce" in n else "", n)
    return n[:96]


def main(path, out):
    data = json.load(open(path))
    ev = sorted((e for e in data.get("traceEvents", []) if e.get("cat") == "kernel" and "ts" in e), key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "normalize_cl_kernel" in e["name"]]
    step = ev[starts[-2]:starts[-1]]
    span = max(e["ts"] + e["dur"] for e in step) - step[0]["ts"]
    rows = collections.OrderedDict()
    for e in step:
        r = rows.setdefault(short(e["name"]), [0, 0.0])
        r[0] += 1
        r[1] += e["dur"]
    tot = sum(v[1] for v in rows.values())
    ours = sum(v[1] for k, v in rows.items() if k.startswith("b200::"))
    nours = sum(v[0] for k, v in rows.items() if k.startswith("b200::"))
    lines = ["# Launch list: one ResNet-50 bf16 training step (batch 32, 1 GPU, CUDA-graph replay), round 2", "",
             f"source: CUPTI kernel records of the last complete step (`bench.py --trace_dir`, `tools/trace_launch_list.py`): {len(step)} kernels, step span "
             f"{span:.0f} us under the tracer (4720 us without), summed kernel time {tot:.0f} us.  Shares only - a traced run is not a timing source.", "",
             f"Kernels of this framework: {nours} launches, {ours:.0f} us ({100 * ours / tot:.1f} %); library (cuDNN / cuBLASLt `nvjet` / ATen): "
             f"{len(step) - nours} launches, {tot - ours:.0f} us.", "",
             "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        tag = " **(ours)**" if k.startswith("b200::") else ""
        lines.append(f"| `{k}`{tag} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
