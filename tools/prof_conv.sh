#!/bin/bash
# ncu --set full captures of the convolution kernels (one GPU).  Usage: bash tools/prof_conv.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
prof() {  # tag, kernel regex, args...
  local tag=$1; shift; local rx=$1; shift
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o gpurun_out/prof_$tag python bench/conv_one.py "$@" > gpurun_out/prof_$tag.log 2>&1
  echo "$tag rc=$?"
}
prof conv_l1c1_f conv_tap_gemm --layer "l1.c1 1x1" --op f --bn 64
prof conv_l1c2_f conv_tap_gemm --layer "l1.c2 3x3" --op f --mode 1 --bn 64
prof conv_l1c2_f_halo conv_tap_gemm --layer "l1.c2 3x3" --op f --mode 2 --bn 64
prof conv_l4c1_f conv_tap_gemm --layer "l4.c1 1x1" --op f --bn 64
prof conv_l1c3_w conv_wgrad_kernel --layer "l1.c3 1x1" --op w
prof conv_l3c2_w conv_wgrad_kernel --layer "l3.c2 3x3" --op w
