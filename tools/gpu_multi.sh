#!/bin/bash
# Multi-GPU run (gpurun --gpus N): peer-memory transport tests, allreduce sweep vs NCCL, DDP benches.
cd "$(dirname "$0")/.."
N=${1:-2}
STEPS=${2:-30}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
echo "== multi tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $O/test_multi_$N.log 2>&1; echo "multi tests rc=$?" | tee -a $O/summary_$N.txt
echo "== sweep"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 \
   bench/allreduce_sweep.py --max_mb 256 --out $O/sweep_$N.json > $O/sweep_$N.log 2>&1; echo "sweep rc=$?" | tee -a $O/summary_$N.txt
for impl in ours reference; do
  echo "== bench $impl N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 \
     bench.py --impl $impl --gpus $N --steps $STEPS --warmup 8 > $O/bench_${impl}_$N.json 2> $O/bench_${impl}_$N.err
  echo "bench $impl rc=$?" | tee -a $O/summary_$N.txt
done
echo "== bench ours nccl transport N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29603 \
   bench.py --gpus $N --steps $STEPS --warmup 8 --backend nccl --skip_e2e > $O/bench_ours_nccl_$N.json 2> $O/bench_ours_nccl_$N.err
echo "bench ours-nccl rc=$?" | tee -a $O/summary_$N.txt
tail -n 15 $O/test_multi_$N.log
tail -n 25 $O/sweep_$N.log
cat $O/bench_ours_$N.json $O/bench_reference_$N.json $O/bench_ours_nccl_$N.json
tail -n 8 $O/bench_ours_$N.err
