#!/bin/bash
# compute-sanitizer over the single-GPU kernels (the reference has no race detection of any kind).
# memcheck on everything, racecheck + synccheck on the kernels that use shared memory / mbarriers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_small_linear_fwd_bwd or test_mse_fused or test_cross_entropy_fused or test_layernorm or test_fused_sgd or test_normalize'
for tool in memcheck racecheck synccheck; do
  echo "== $tool (elementwise / reduction kernels)"
  timeout 1200 $SAN --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$SEL" > $O/sanitize_${tool}.log 2>&1
  echo "$tool kernels rc=$?" | tee -a $O/sanitize_summary.txt
done
echo "== memcheck + racecheck (BatchNorm / pooling kernels)"
for tool in memcheck racecheck; do
  timeout 1200 $SAN --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_batchnorm.py -m gpu -q \
      -k "shape2 or shape4 or maxpool" > $O/sanitize_bn_${tool}.log 2>&1
  echo "$tool bn rc=$?" | tee -a $O/sanitize_summary.txt
done
echo "== memcheck (tcgen05 GEMM, small shapes)"
timeout 1200 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "test_gemm_operand_layouts or test_gemm_epilogues" > $O/sanitize_gemm_memcheck.log 2>&1
echo "memcheck gemm rc=$?" | tee -a $O/sanitize_summary.txt
grep -h "ERROR SUMMARY\|passed\|failed" $O/sanitize_*.log
