#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/test_conv.log 2>&1; echo "conv tests rc=$?"; tail -n 12 gpurun_out/test_conv.log
B="python bench.py --gpus 1 --steps 40 --warmup 8 --skip_e2e"
run() {
  local tag=$1; shift
  env "$@" timeout 200 $B > gpurun_out/step_$tag.json 2> gpurun_out/step_$tag.err; echo "$tag rc=$?"
  python - gpurun_out/step_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"   {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step launches/step={d.get('native_launches_per_step')}")
except Exception as e:
    print("   no result:", e)
PY
}
run fused_auto B200DDP_CONV=native B200DDP_CONV_WGRAD=auto B200DDP_BLOCK_FUSE=1
run fused_wlib B200DDP_CONV=native B200DDP_CONV_WGRAD=lib B200DDP_BLOCK_FUSE=1
run unfused_wlib B200DDP_CONV=native B200DDP_CONV_WGRAD=lib B200DDP_BLOCK_FUSE=0
run lib B200DDP_CONV=lib B200DDP_CONV_WGRAD=lib
