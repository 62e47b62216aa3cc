#!/bin/bash
# First GPU call of the next round: validate and measure every opt-in path prepared without a GPU, one JSON line each.
#   gpurun --timeout 1500 -- bash tools/round2_ablation.sh
# Output: gpurun_out/ablation_*.json (+ .err), gpurun_out/kernels_groupm.log, gpurun_out/test_optin.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --gpus 1 --steps 60 --warmup 10 --skip_e2e"
echo "== opt-in tests"; B200DDP_TEST_OPTIN=1 timeout 600 python -m pytest tests/test_gpu_zz_optin.py tests/test_gpu_zz_gemm_raster.py -m gpu -q > $O/test_optin.log 2>&1; echo "rc=$?"; tail -n 5 $O/test_optin.log
run() {  # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 400 $B > $O/ablation_$name.json 2> $O/ablation_$name.err; echo "rc=$?"
  python - "$O/ablation_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"   {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  opt_in={d['config'].get('opt_in')}")
except Exception as e:
    print("   no result:", e)
PY
}
run base A=0
run base_again A=0
run stem_pad8 B200DDP_STEM_PAD=8
run conv1x1_tc B200DDP_CONV1X1_TC=1
run conv1x1_tc_tmastore B200DDP_CONV1X1_TC=1 B200DDP_GEMM_TMA_STORE=1
run conv1x1_tc_tmastore_bnfuse B200DDP_CONV1X1_TC=1 B200DDP_GEMM_TMA_STORE=1 B200DDP_CONV_BN_FUSE=1
run conv1x1_tc_group8 B200DDP_CONV1X1_TC=1 B200DDP_GEMM_GROUP_M=8
run stem_pad8_conv1x1 B200DDP_STEM_PAD=8 B200DDP_CONV1X1_TC=1
run bn_fused B200DDP_BN_FUSED=1
run bn_pdl B200DDP_PDL=1
run conv3x3_tc B200DDP_CONV3X3_TC=1
run all_conv_tc B200DDP_STEM_PAD=8 B200DDP_CONV1X1_TC=1 B200DDP_GEMM_TMA_STORE=1 B200DDP_CONV_BN_FUSE=1 B200DDP_CONV3X3_TC=1
echo "== conv3x3 draft kernel vs cuDNN (time)"; timeout 300 python bench/conv_bench.py > $O/conv_bench.log 2>&1; echo "rc=$?"; cat $O/conv_bench.log
echo "== GEMM rasterisation sweep"; timeout 600 python bench/kernel_bench.py --only gemm --out $O/kernels_groupm.json > $O/kernels_groupm.log 2>&1; echo "rc=$?"; grep -i "gemm" $O/kernels_groupm.log
echo "== launch list with the opt-ins on (who replaced whom)"
B200DDP_STEM_PAD=8 B200DDP_CONV1X1_TC=1 timeout 400 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum \
    --profile-from-start off --csv --log-file $O/launches_optin.csv python bench.py --steps 2 --warmup 6 --skip_e2e --profile_range \
    > $O/launches_optin_bench.json 2> $O/launches_optin.err; echo "rc=$?"
