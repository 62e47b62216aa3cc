#!/bin/bash
# Per-kernel device time of one ResNet-50 step (graph replay), for the native conv path and the library conv path.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in native lib; do
  B200DDP_CONV=$cfg B200DDP_CONV_WGRAD=${WGRAD:-auto} timeout 400 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum \
    --profile-from-start off --csv --log-file gpurun_out/launches_$cfg.csv python bench.py --steps 2 --warmup 6 --skip_e2e --profile_range \
    > gpurun_out/launches_${cfg}_bench.json 2> gpurun_out/launches_$cfg.err; echo "$cfg rc=$?"
done
