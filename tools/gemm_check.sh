#!/bin/bash
# 1-GPU job: fused BatchNorm + 2-CTA GEMM correctness (own timeouts: a hang must not eat the lease), then timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_batchnorm.py -m gpu -q -x > $O/test_bn.log 2>&1; echo "bn tests rc=$?"
tail -n 8 $O/test_bn.log
timeout 300 python -m pytest tests/test_gpu_gemm_2cta.py -m gpu -q -x > $O/test_gemm_2cta.log 2>&1; echo "2cta tests rc=$?"
tail -n 8 $O/test_gemm_2cta.log
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -q > $O/test_gemm_all.log 2>&1; echo "all gemm+kernel tests rc=$?"
tail -n 3 $O/test_gemm_all.log
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 8 > $O/bench_ours.json 2> $O/bench_ours.err; echo "bench rc=$?"
cat $O/bench_ours.json; tail -n 3 $O/bench_ours.err
timeout 600 python bench/kernel_bench.py --only gemm --out $O/kernels_gemm.json > $O/kernels_gemm.log 2>&1; echo "gemm bench rc=$?"
cat $O/kernels_gemm.log
