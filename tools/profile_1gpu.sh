#!/bin/bash
# Kernel timings vs the measured roofline + ncu captures (1 GPU).  Results land in gpurun_out/; summaries are
# copied to profiles/ by tools/summarize_profiles.py on the CPU box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
echo "== kernel bench"; timeout 900 python bench/kernel_bench.py --out $O/kernels.json > $O/kernels.log 2>&1; echo "kernel bench rc=$?"
echo "== launch list of one bench step (eager, so every kernel is visible)"
timeout 600 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 6 --no_graph --skip_e2e --profile_range > $O/launches_bench.json 2> $O/launches.err; echo "launch list rc=$?"
echo "== launch list of one bench step (CUDA graph replay)"
timeout 600 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/launches_graph.csv \
    python bench.py --steps 2 --warmup 6 --skip_e2e --profile_range > $O/launches_graph_bench.json 2> $O/launches_graph.err; echo "graph launch list rc=$?"
echo "== ncu full: tcgen05 GEMM"
timeout 600 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -o $O/prof_gemm -f \
    python bench/kernel_bench.py --only gemm --iters 1 > $O/prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
echo "== ncu full: fused SGD"
timeout 600 $NCU --set full --import-source on -k regex:multi_sgd_kernel -s 3 -c 1 -o $O/prof_sgd -f \
    python bench/kernel_bench.py --only sgd --iters 1 > $O/prof_sgd.log 2>&1; echo "ncu sgd rc=$?"
echo "== ncu full: layernorm bwd + xent"
timeout 600 $NCU --set full --import-source on -k regex:layernorm_bwd_fast_kernel -s 3 -c 1 -o $O/prof_lnbwd -f \
    python bench/kernel_bench.py --only ln --iters 1 > $O/prof_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:xent_fwd_bwd_kernel -s 3 -c 1 -o $O/prof_xent -f \
    python bench/kernel_bench.py --only xent --iters 1 > $O/prof_xent.log 2>&1; echo "ncu xent rc=$?"
cat $O/kernels.log
ls -la $O/*.ncu-rep
