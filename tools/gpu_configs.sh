#!/bin/bash
# BASELINE.json configs 3 and 4 on N GPUs: ours vs stock torch DDP (+NCCL, cuBLAS/cuDNN) on the same box.
#   3: BERT-base DDP bf16, seq 512, bucket_cap_mb sweep
#   4: ResNet-152 DDP with gradient_as_bucket_view + find_unused_parameters
cd "$(dirname "$0")/.."
N=${1:-8}
CAPS=${2:-"1 25 256"}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/configs_$N.jsonl
echo "== multi tests"; timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests rc=$?"; tail -n 4 $O/test_multi_$N.log
launch() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 6 --skip_e2e "$@"; }
for cap in $CAPS; do
  for impl in ours stock; do
    launch --impl $impl --model bert-base --bucket_cap_mb $cap 2> $O/cfg_bert_${impl}_${cap}_$N.err | grep '^{' | tee -a $O/configs_$N.jsonl | python -c "import sys,json; [print('bert cap', $cap, d['impl'], round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms') for d in map(json.loads, sys.stdin)]"
  done
done
for impl in ours stock; do
  launch --impl $impl --model resnet152 --gradient_as_bucket_view --find_unused_parameters 2> $O/cfg_rn152_${impl}_$N.err | grep '^{' | tee -a $O/configs_$N.jsonl | python -c "import sys,json; [print('resnet152 gabv+unused', d['impl'], round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms') for d in map(json.loads, sys.stdin)]"
done

tail -n 5 $O/cfg_bert_ours_*_$N.err | cut -c1-300
tail -n 5 $O/cfg_rn152_ours_$N.err | cut -c1-300
