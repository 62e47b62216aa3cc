#!/usr/bin/env python
"""Run ONE of our convolution kernels a few times (for `ncu -k regex:conv_ ...`).

  python bench/conv_one.py --layer "l1.c1 1x1" --op f [--mode -1 --bn 0] [--split 0 --tm 0 --tn 0] [--iters 3]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bench"))


def main():
    from conv_layers import LAYERS
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", required=True)
    ap.add_argument("--op", default="f", choices=["f", "d", "w"])
    ap.add_argument("--mode", type=int, default=-1)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--tm", type=int, default=0)
    ap.add_argument("--tn", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    from b200ddp import _ext
    C = _ext.get()
    (name, ci, co, k, s, h, cnt) = [l for l in LAYERS if l[0] == a.layer][0]
    dev = torch.device("cuda", 0)
    pad = k // 2
    x = torch.randn(a.batch, ci, h, h, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(a.batch, co, h, h, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(a.iters):
        flush.fill_(1)
        if a.op == "f":
            C.conv_fprop(x, w, 1, pad, a.mode, a.bn, 0, False)
        elif a.op == "d":
            C.conv_dgrad(dy, w, 1, pad, a.mode, a.bn, 0)[0]
        else:
            C.conv_wgrad(dy, x, k, 1, pad, a.split, a.tm, a.tn)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
