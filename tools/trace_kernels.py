#!/usr/bin/env python
"""Aggregate kernel durations of the last complete step of a `bench.py --trace_dir` CUPTI trace by kernel family."""
import collections, json, re, sys
def main(path):
    data = json.load(open(path))
    ev = [e for e in data.get("traceEvents", []) if e.get("cat") == "kernel" and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "normalize_cl_kernel" in e["name"]]
    a, b = starts[-2], starts[-1]
    step = ev[a:b]
    span = max(e["ts"] + e["dur"] for e in step) - step[0]["ts"]
    groups = collections.OrderedDict()
    for e in step:
        k = e["name"]
        if "conv_tap_gemm" in k:
            m = re.search(r"conv_tap_gemm_kernel<(?:\(int\))?(\d+), (?:\(bool\))?(\w+), (?:\(int\))?(\d+)>", k)
            g = f"ours conv bn={m.group(1)} {'dgrad' if m.group(2) in ('1','true') else 'fprop'} epi={m.group(3)}" if m else "ours conv"
        elif "conv_wgrad_kernel" in k: g = "ours wgrad"
        elif "conv_wgrad_reduce" in k: g = "ours wgrad reduce"
        elif "implicit_gemm" in k or "xmma" in k or "cudnn" in k:
            g = "cudnn " + ("wgrad" if "wgrad" in k else "dgrad" if "dgrad" in k else "fprop")
        else:
            g = re.sub(r"\(.*", "", k).replace("void ", "").replace("b200::<unnamed>::", "").replace("b200::(anonymous namespace)::", "")
            g = re.sub(r"<.*", "", g)[:60]
        x = groups.setdefault(g, [0, 0.0]); x[0] += 1; x[1] += e["dur"]
    tot = sum(v[1] for v in groups.values())
    print(f"{path}: step span {span:.0f} us, {len(step)} kernels, summed kernel time {tot:.0f} us")
    for k, v in sorted(groups.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"  {v[0]:4d} {v[1]:8.1f}  {v[1]/v[0]:6.1f}  {k}")
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
