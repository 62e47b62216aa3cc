#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/bert_caps_$N.jsonl
for cap in 1 256; do
  for impl in ours stock; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 \
       bench.py --impl $impl --model bert-base --gpus $N --steps 20 --warmup 6 --skip_e2e --bucket_cap_mb $cap 2> $O/bertcap_${impl}_${cap}_$N.err | grep '^{' | tee -a $O/bert_caps_$N.jsonl | python -c "import sys,json; [print('bert cap', $cap, d['impl'], round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', d['config'].get('ddp',{}).get('buckets')) for d in map(json.loads, sys.stdin)]"
  done
done
