#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
export B200DDP_TEST_WORLD=$N
for t in test_peer_collectives test_ddp_foo_matches_stock_ddp test_ddp_bf16_and_unused_parameters; do
  timeout 170 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k $t > $O/dbg_${t}_$N.log 2>&1; echo "$t rc=$?"
  grep -E "collectives\]|passed|failed|Error|error|assert" $O/dbg_${t}_$N.log | tail -n 12 | cut -c1-220
done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 \
   bench/allreduce_sweep.py --max_mb 64 --out $O/sweep_$N.json > $O/sweep_$N.log 2>&1; echo "sweep rc=$?"
tail -n 10 $O/sweep_$N.log | cut -c1-280
