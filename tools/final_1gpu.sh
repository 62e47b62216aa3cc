#!/bin/bash
# Round-end 1-GPU evidence run: GPU test tier, both bench arms, kernel roofline, launch list, ncu captures, sanitizer subset.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
SAN=/usr/local/cuda/bin/compute-sanitizer
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -x -q > $O/test_gpu_all.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 $O/test_gpu_all.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/smoke.log
echo "== bench ours"; timeout 400 python bench.py --gpus 1 --steps 50 --warmup 8 > $O/bench_ours.json 2> $O/bench_ours.err; echo "rc=$?"; cat $O/bench_ours.json
echo "== bench reference"; timeout 400 python bench.py --impl reference --gpus 1 --steps 50 --warmup 8 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cat $O/bench_ref.json
echo "== bench foo (the reference's own workload), both arms"
timeout 200 python bench.py --model foo --gpus 1 --steps 200 --warmup 20 > $O/bench_foo_ours.json 2> $O/bench_foo_ours.err; echo "rc=$?"; cat $O/bench_foo_ours.json
timeout 200 python bench.py --impl reference --model foo --gpus 1 --steps 200 --warmup 20 > $O/bench_foo_ref.json 2> $O/bench_foo_ref.err; echo "rc=$?"; cat $O/bench_foo_ref.json
echo "== kernel bench"; timeout 600 python bench/kernel_bench.py --out $O/kernels.json > $O/kernels.log 2>&1; echo "rc=$?"; cat $O/kernels.log
echo "== launch list (graph replay)"
timeout 400 $NCU --cache-control none --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/launches_graph.csv \
    python bench.py --steps 2 --warmup 6 --skip_e2e --profile_range > $O/launches_graph_bench.json 2> $O/launches_graph.err; echo "rc=$?"
echo "== ncu full captures"
timeout 300 $NCU --set full --import-source on -k regex:gemm_bf16_2cta_kernel -s 3 -c 1 -o $O/prof_gemm2cta -f python bench/kernel_bench.py --only gemm --iters 1 > $O/prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:bn_bwd_reduce_kernel -s 8 -c 1 -o $O/prof_bnbwd -f python bench/kernel_bench.py --only bn --iters 1 > $O/prof_bn.log 2>&1; echo "ncu bn rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:bn_stats_kernel -s 8 -c 1 -o $O/prof_bnstats -f python bench/kernel_bench.py --only bn --iters 1 > $O/prof_bn2.log 2>&1; echo "ncu bn stats rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:xent_fwd_bwd_smem_kernel -s 3 -c 1 -o $O/prof_xent -f python bench/kernel_bench.py --only xent --iters 1 > $O/prof_xent.log 2>&1; echo "ncu xent rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:multi_sgd_kernel -s 3 -c 1 -o $O/prof_sgd -f python bench/kernel_bench.py --only sgd --iters 1 > $O/prof_sgd.log 2>&1; echo "ncu sgd rc=$?"
echo "== sanitizer (memcheck + racecheck on the reduction / elementwise kernels, memcheck on small GEMMs)"
SEL='test_small_linear_fwd_bwd or test_mse_fused or test_layernorm or test_fused_sgd or test_normalize or test_gelu'
timeout 400 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$SEL" > $O/sanitize_memcheck.log 2>&1; echo "memcheck kernels rc=$?" | tee -a $O/sanitize_summary.txt
timeout 400 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "test_mse_fused or test_layernorm or test_small_linear_fwd_bwd" > $O/sanitize_racecheck.log 2>&1; echo "racecheck kernels rc=$?" | tee -a $O/sanitize_summary.txt
timeout 400 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_batchnorm.py -m gpu -q -k "shape2 or shape4 or maxpool" > $O/sanitize_bn_memcheck.log 2>&1; echo "memcheck bn/pool rc=$?" | tee -a $O/sanitize_summary.txt
timeout 400 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "test_gemm_epilogues or test_gemm_fp32_accumulate_into" > $O/sanitize_gemm_memcheck.log 2>&1; echo "memcheck gemm rc=$?" | tee -a $O/sanitize_summary.txt
grep -h "ERROR SUMMARY\|passed\|failed" $O/sanitize_*.log | sort | uniq -c
ls -la $O/*.ncu-rep
