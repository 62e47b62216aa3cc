#!/usr/bin/env python
"""Per-layer convolution table for ResNet-50 (batch 32, bf16, channels_last): cuDNN fprop / dgrad / wgrad time per
distinct layer shape (the bar), next to this framework's tcgen05 kernels for the shapes they cover.

  python bench/conv_layers.py [--iters 20] [--batch 32] [--out gpurun_out/conv_layers.json]

CUDA events, 3 warm-up launches, a 256 MB write between timed launches (L2 flush), median of `iters`.
The reference's hot path these replace: the model's forward / backward at /root/reference/ddp.py:221,231."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, C_in, C_out, kernel, stride, H_in, count in ResNet-50)
LAYERS = [
    ("stem7x7s2", 3, 64, 7, 2, 224, 1),
    ("l1.c1a 1x1", 64, 64, 1, 1, 56, 1),
    ("l1.c1 1x1", 256, 64, 1, 1, 56, 2),
    ("l1.c2 3x3", 64, 64, 3, 1, 56, 3),
    ("l1.c3 1x1", 64, 256, 1, 1, 56, 4),      # 3 x conv3 + the stride-1 downsample
    ("l2.c1a 1x1", 256, 128, 1, 1, 56, 1),
    ("l2.c2a 3x3s2", 128, 128, 3, 2, 56, 1),
    ("l2.ds 1x1s2", 256, 512, 1, 2, 56, 1),
    ("l2.c1 1x1", 512, 128, 1, 1, 28, 3),
    ("l2.c2 3x3", 128, 128, 3, 1, 28, 3),
    ("l2.c3 1x1", 128, 512, 1, 1, 28, 4),
    ("l3.c1a 1x1", 512, 256, 1, 1, 28, 1),
    ("l3.c2a 3x3s2", 256, 256, 3, 2, 28, 1),
    ("l3.ds 1x1s2", 512, 1024, 1, 2, 28, 1),
    ("l3.c1 1x1", 1024, 256, 1, 1, 14, 5),
    ("l3.c2 3x3", 256, 256, 3, 1, 14, 5),
    ("l3.c3 1x1", 256, 1024, 1, 1, 14, 6),
    ("l4.c1a 1x1", 1024, 512, 1, 1, 14, 1),
    ("l4.c2a 3x3s2", 512, 512, 3, 2, 14, 1),
    ("l4.ds 1x1s2", 1024, 2048, 1, 2, 14, 1),
    ("l4.c1 1x1", 2048, 512, 1, 1, 7, 2),
    ("l4.c2 3x3", 512, 512, 3, 1, 7, 2),
    ("l4.c3 1x1", 512, 2048, 1, 1, 7, 3),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", type=str, default=None)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    from b200ddp import _ext
    C = _ext.get()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts)

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

    rows = []
    tot = {"lib": 0.0, "ours": 0.0}
    n = args.batch
    hdr = f"{'layer':14s} {'x':>3s} | {'fprop lib':>9s} {'ours':>7s} | {'dgrad lib':>9s} {'ours':>7s} | {'wgrad lib':>9s} {'ours':>7s} | relerr f/d/w"
    print(hdr, flush=True)
    for (name, ci, co, k, s, h, cnt) in LAYERS:
        if args.only and args.only not in name:
            continue
        pad = k // 2
        ho = (h + 2 * pad - k) // s + 1
        x = torch.randn(n, ci, h, h, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(co, ci, k, k, device=dev) * (1.0 / (ci * k * k) ** 0.5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dy = torch.randn(n, co, ho, ho, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

        def bwd(mask):
            return torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, mask)

        lib_f = timeit(lambda: F.conv2d(x, w, None, s, pad))
        lib_d = timeit(lambda: bwd([True, False, False])) if ci > 3 else float("nan")
        lib_w = timeit(lambda: bwd([False, True, False]))
        row = {"layer": name, "count": cnt, "cin": ci, "cout": co, "k": k, "stride": s, "h": h, "lib_fprop_us": lib_f, "lib_dgrad_us": lib_d,
               "lib_wgrad_us": lib_w}
        our_f = our_d = our_w = float("nan")
        errs = ""
        try:
            if hasattr(C, "conv_fprop"):
                y_ref = F.conv2d(x, w, None, s, pad)
                our_f = timeit(lambda: C.conv_fprop(x, w, s, pad))
                errs += f"{rel(C.conv_fprop(x, w, s, pad), y_ref):.1e}"
                if ci > 3 and hasattr(C, "conv_dgrad"):
                    dx_ref = bwd([True, False, False])[0]
                    our_d = timeit(lambda: C.conv_dgrad(dy, w, s, pad, h, h))
                    errs += f"/{rel(C.conv_dgrad(dy, w, s, pad, h, h), dx_ref):.1e}"
                if hasattr(C, "conv_wgrad"):
                    dw_ref = bwd([False, True, False])[1]
                    our_w = timeit(lambda: C.conv_wgrad(dy, x, k, s, pad))
                    errs += f"/{rel(C.conv_wgrad(dy, x, k, s, pad), dw_ref):.1e}"
            elif k == 1 and s == 1:
                x2 = x.permute(0, 2, 3, 1).reshape(-1, ci)
                dy2 = dy.permute(0, 2, 3, 1).reshape(-1, co)
                w2 = w.reshape(co, ci)
                our_f = timeit(lambda: C.gemm_nt(x2, w2, None, 0, None))
                our_d = timeit(lambda: C.gemm(dy2, w2, None, False, True, 0, False, None))
                our_w = timeit(lambda: C.gemm(dy2, x2, None, True, True, 0, False, None))
                y_ref = F.conv2d(x, w).permute(0, 2, 3, 1).reshape(-1, co)
                g = bwd([True, True, False])
                errs = (f"{rel(C.gemm_nt(x2, w2, None, 0, None), y_ref):.1e}/"
                        f"{rel(C.gemm(dy2, w2, None, False, True, 0, False, None), g[0].permute(0, 2, 3, 1).reshape(-1, ci)):.1e}/"
                        f"{rel(C.gemm(dy2, x2, None, True, True, 0, False, None), g[1].reshape(co, ci)):.1e}")
            elif k == 3 and s == 1 and ci % 64 == 0:
                our_f = timeit(lambda: C.conv3x3_fwd(x, w))
                errs = f"{rel(C.conv3x3_fwd(x, w), F.conv2d(x, w, padding=1)):.1e}"
        except Exception as exc:  # keep the table going; a failing shape is a finding, not a crash
            errs = f"ERR {type(exc).__name__}: {str(exc)[:80]}"
        row.update({"ours_fprop_us": our_f, "ours_dgrad_us": our_d, "ours_wgrad_us": our_w, "relerr": errs})
        rows.append(row)
        for lib, ours in ((lib_f, our_f), (lib_d, our_d), (lib_w, our_w)):
            if lib == lib:
                tot["lib"] += lib * cnt
                tot["ours"] += (ours if ours == ours else lib) * cnt
        print(f"{name:14s} {cnt:3d} | {lib_f:9.1f} {our_f:7.1f} | {lib_d:9.1f} {our_d:7.1f} | {lib_w:9.1f} {our_w:7.1f} | {errs}", flush=True)
    print(f"sum over the network (count-weighted, library time where we have no kernel): cuDNN {tot['lib']:.0f} us, ours {tot['ours']:.0f} us", flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"batch": n, "rows": rows, "total_us": tot}, f, indent=1)


if __name__ == "__main__":
    main()
