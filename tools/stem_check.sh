#!/bin/bash
# Stem kernels: numerics tests, per-op timing against the library, and the training step with / without them (1 GPU).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k stem > $O/test_stem.log 2>&1; echo "stem tests rc=$? : $(tail -n 1 $O/test_stem.log)"
grep -n "Error\|error:\|assert" $O/test_stem.log | head -12
timeout 120 python bench/stem_bench.py --out $O/stem_bench.json 2>&1 | tail -30 | cut -c1-250
for mode in native lib; do
  B200DDP_STEM=$mode timeout 200 python bench.py --steps 60 --warmup 8 --skip_e2e > $O/step_stem_$mode.json 2> $O/step_stem_$mode.err
  echo "step stem=$mode rc=$? $(python -c "import json; d=json.loads([l for l in open('$O/step_stem_$mode.json') if l.startswith('{')][-1]); print(round(d['ms_per_step'],4),'ms', d.get('native_launches_per_step'))" 2>&1)"
done
tail -n 3 $O/step_stem_native.err
