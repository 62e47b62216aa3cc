#!/bin/bash
cd "$(dirname "$0")/.."
N=2
mkdir -p gpurun_out
O=gpurun_out
echo "== multi tests"; timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests rc=$?"; tail -n 3 $O/test_multi_$N.log
for impl in ours reference; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
     bench.py --impl $impl --gpus $N --steps 40 --warmup 8 > $O/bench_${impl}_$N.json 2> $O/bench_${impl}_$N.err
  echo "bench $impl rc=$?"; grep '^{' $O/bench_${impl}_$N.json | cut -c1-420
done
for impl in ours stock; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 \
     bench.py --impl $impl --model bert-base --gpus $N --steps 20 --warmup 6 --skip_e2e --bucket_cap_mb 25 > $O/bert_${impl}_$N.json 2> $O/bert_${impl}_$N.err
  echo "bert $impl rc=$?"; grep '^{' $O/bert_${impl}_$N.json | cut -c1-300
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 \
   bench.py --model foo --gpus $N --steps 200 --warmup 20 > $O/bench_foo_ours_$N.json 2> $O/bench_foo_ours_$N.err; echo "foo ours rc=$?"; grep '^{' $O/bench_foo_ours_$N.json | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29644 \
   bench.py --impl reference --model foo --gpus $N --steps 200 --warmup 20 > $O/bench_foo_ref_$N.json 2> $O/bench_foo_ref_$N.err; echo "foo ref rc=$?"; grep '^{' $O/bench_foo_ref_$N.json | cut -c1-300
tail -n 4 $O/bert_ours_$N.err | cut -c1-300
