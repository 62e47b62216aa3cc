#!/bin/bash
# GEMM-focused 1-GPU job: 2-CTA correctness (own timeout: a hang must not eat the lease) + timings for both modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_2cta.py -m gpu -q -x > $O/test_gemm_2cta.log 2>&1; echo "2cta tests rc=$?"
tail -n 6 $O/test_gemm_2cta.log
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -q > $O/test_gemm_all.log 2>&1; echo "all gemm+kernel tests rc=$?"
tail -n 3 $O/test_gemm_all.log
timeout 600 python bench/kernel_bench.py --only gemm --out $O/kernels_gemm.json > $O/kernels_gemm.log 2>&1; echo "gemm bench rc=$?"
cat $O/kernels_gemm.log
