#!/usr/bin/env python
"""Cycle breakdown of CTA 0's producer / MMA issuer / epilogue in conv_tap_gemm_kernel (diagnostics)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bench"))
from conv_layers import LAYERS
from b200ddp import _ext
C = _ext.get()
dev = torch.device("cuda", 0)
names = ["prod wait A-empty", "prod issue A", "prod wait B-empty", "prod issue B", "mma wait tmem-empty", "mma wait full", "mma issue", "mma commit",
         "epi wait tmem-full", "epi work", "kernel total"]
for (layer, mode, bn) in [("l1.c1 1x1", -1, 64), ("l4.c1 1x1", -1, 64), ("l4.c1 1x1", -1, 256), ("l1.c2 3x3", 1, 64), ("l1.c2 3x3", 2, 64), ("l3.c2 3x3", 1, 128)]:
    (name, ci, co, k, s, h, cnt) = [l for l in LAYERS if l[0] == layer][0]
    x = torch.randn(32, ci, h, h, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dbg = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(3):
        C.conv_fprop(x, w, 1, k // 2, mode, bn, 0, False, dbg)
    torch.cuda.synchronize()
    d = dbg.tolist()
    kblocks = (ci // 64) * k * k
    print(f"{layer} mode={mode} bn={bn}: k-blocks/tile={kblocks}")
    for n, v in zip(names, d):
        print(f"    {n:22s} {v:9d} cyc")
