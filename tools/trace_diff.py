#!/usr/bin/env python
"""Align the compute kernels of two `bench.py --trace_dir` traces (same model, e.g. with and without gradient communication)
and report where the step got longer: per-kernel duration deltas and inter-kernel gap deltas, aggregated by kernel family."""
import collections, json, re, sys

def load(path):
    data = json.load(open(path))
    ev = [e for e in data.get("traceEvents", []) if e.get("cat") == "kernel" and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "normalize_cl_kernel" in e["name"]]
    step = ev[starts[-2]:starts[-1]]
    comm = lambda e: "bucket_allreduce" in e["name"] or "peer_broadcast" in e["name"] or "nccl" in e["name"].lower()   # noqa: E731
    return [e for e in step if not comm(e)], [e for e in step if comm(e)]

def fam(name):
    n = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::<unnamed>::", "").replace("b200::(anonymous namespace)::", "")
    if "implicit_gemm" in n or "xmma" in n or "cudnn" in n:
        return "cudnn " + ("wgrad" if "wgrad" in n else "dgrad" if "dgrad" in n else "fprop")
    return re.sub(r"<.*", "", n)[:48]

def main(a_path, b_path):
    a, _ = load(a_path)
    b, bc = load(b_path)
    print(f"A (reference) {a_path}: {len(a)} compute kernels, span {a[-1]['ts'] + a[-1]['dur'] - a[0]['ts']:.0f} us")
    print(f"B            {b_path}: {len(b)} compute kernels, span {b[-1]['ts'] + b[-1]['dur'] - b[0]['ts']:.0f} us, {len(bc)} comm kernels")
    if len(a) != len(b):
        print("kernel counts differ; aligning the common prefix")
    n = min(len(a), len(b))
    dur = collections.OrderedDict(); gap = collections.OrderedDict()
    tot_d = tot_g = 0.0
    windows = [(e["ts"], e["ts"] + e["dur"]) for e in bc]
    in_comm = out_comm = 0.0
    for i in range(n):
        d = b[i]["dur"] - a[i]["dur"]
        f = fam(a[i]["name"])
        x = dur.setdefault(f, [0, 0.0, 0.0]); x[0] += 1; x[1] += d; x[2] += a[i]["dur"]
        tot_d += d
        mid = b[i]["ts"] + b[i]["dur"] / 2
        if any(lo <= mid <= hi for lo, hi in windows): in_comm += d
        else: out_comm += d
        if i:
            ga = a[i]["ts"] - (a[i - 1]["ts"] + a[i - 1]["dur"]); gb = b[i]["ts"] - (b[i - 1]["ts"] + b[i - 1]["dur"])
            y = gap.setdefault(f, [0, 0.0]); y[0] += 1; y[1] += gb - ga
            tot_g += gb - ga
    print(f"sum of kernel-duration deltas {tot_d:+.0f} us (while a comm kernel is running: {in_comm:+.0f} us, otherwise {out_comm:+.0f} us); sum of gap deltas {tot_g:+.0f} us")
    print("duration delta by family (count, delta us, reference us):")
    for k, v in sorted(dur.items(), key=lambda kv: -abs(kv[1][1]))[:14]:
        print(f"  {v[0]:4d} {v[1]:+8.1f} {v[2]:9.1f}  {k}")
    print("gap-before delta by family:")
    for k, v in sorted(gap.items(), key=lambda kv: -abs(kv[1][1]))[:10]:
        print(f"  {v[0]:4d} {v[1]:+8.1f}  {k}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
