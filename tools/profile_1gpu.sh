#!/bin/bash
# Kernel timings vs the measured roofline + ncu launch lists / captures (1 GPU).
cd "$(dirname "$0")/.."
MODE=${1:-full}
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
echo "== tests (bn, 2cta, kernels)"; timeout 600 python -m pytest tests/test_gpu_batchnorm.py tests/test_gpu_gemm_2cta.py tests/test_gpu_kernels.py -m gpu -q > $O/test_1gpu.log 2>&1; echo "tests rc=$?"; tail -n 3 $O/test_1gpu.log
echo "== bench"; timeout 600 python bench.py --gpus 1 --steps 40 --warmup 8 > $O/bench_ours.json 2> $O/bench_ours.err; echo "bench rc=$?"; cat $O/bench_ours.json; tail -n 3 $O/bench_ours.err
echo "== kernel bench"; timeout 900 python bench/kernel_bench.py --out $O/kernels.json > $O/kernels.log 2>&1; echo "kernel bench rc=$?"
cat $O/kernels.log
echo "== launch list of one bench step (CUDA graph replay)"
timeout 600 $NCU --cache-control none --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/launches_graph.csv \
    python bench.py --steps 2 --warmup 6 --skip_e2e --profile_range > $O/launches_graph_bench.json 2> $O/launches_graph.err; echo "graph launch list rc=$?"
if [ "$MODE" = "full" ]; then
echo "== ncu full: tcgen05 GEMM (2-CTA)"
timeout 600 $NCU --set full --import-source on -k regex:gemm_bf16_2cta_kernel -s 3 -c 1 -o $O/prof_gemm2cta -f \
    python bench/kernel_bench.py --only gemm --iters 1 > $O/prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
echo "== ncu full: bn kernels"
timeout 600 $NCU --set full --import-source on -k regex:bn_stats_kernel -s 3 -c 1 -o $O/prof_bnstats -f \
    python bench/kernel_bench.py --only bn --iters 1 > $O/prof_bn.log 2>&1; echo "ncu bn stats rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:bn_bwd_reduce_kernel -s 3 -c 1 -o $O/prof_bnbwd -f \
    python bench/kernel_bench.py --only bn --iters 1 > $O/prof_bn2.log 2>&1; echo "ncu bn bwd rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:xent_fwd_bwd_smem_kernel -s 3 -c 1 -o $O/prof_xent -f \
    python bench/kernel_bench.py --only xent --iters 1 > $O/prof_xent.log 2>&1; echo "ncu xent rc=$?"
fi
ls -la $O/*.ncu-rep
