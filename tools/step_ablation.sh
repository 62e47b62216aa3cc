cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/test_conv.log 2>&1; echo "conv tests rc=$?"; tail -n 8 gpurun_out/test_conv.log
B="python bench.py --gpus 1 --steps 40 --warmup 8 --skip_e2e"
for cfg in "native:auto" "native:lib" "native:native" "lib:lib"; do
  c=${cfg%%:*}; w=${cfg##*:}
  B200DDP_CONV=$c B200DDP_CONV_WGRAD=$w timeout 200 $B > gpurun_out/step_conv_${c}_wgrad_${w}.json 2> gpurun_out/step_conv_${c}_wgrad_${w}.err; echo "conv=$c wgrad=$w rc=$?"
  python - gpurun_out/step_conv_${c}_wgrad_${w}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"   {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step launches/step={d.get('native_launches_per_step')}")
except Exception as e:
    print("   no result:", e)
PY
done
