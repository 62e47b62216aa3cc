#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 70 python -m pytest tests -m gpu -q > $O/test_gpu_last.log 2>&1; echo "gpu tests rc=$? : $(tail -n 1 $O/test_gpu_last.log)"
grep -n "^FAILED\|^ERROR" $O/test_gpu_last.log | head -8
timeout 40 python bench/kernel_bench.py --only conv --iters 10 --out $O/kernels_conv.json > $O/kernels_conv.log 2>&1; echo "kernel_bench conv rc=$?"; cut -c1-170 $O/kernels_conv.log | tail -n 22
