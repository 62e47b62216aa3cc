#!/bin/bash
# N-GPU diagnosis: where does the DDP step lose time relative to 1 GPU?  Every stage has a short timeout.
cd "$(dirname "$0")/.."
N=${1:-2}
MODE=${2:-full}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/diag_summary_$N.txt
run() {  # name ENV=.. (bench args in $EXTRA)
  local name=$1; shift
  timeout 150 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 40 --warmup 8 --skip_e2e $EXTRA > $O/diag_${name}_$N.json 2> $O/diag_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$O/diag_${name}_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],3),'ms', d['config'].get('transport'), d['config'].get('ddp',{}).get('algos'), d['config'].get('ddp',{}).get('blocks'))" 2>&1)" | tee -a $O/diag_summary_$N.txt
}
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests rc=$?" | tee -a $O/diag_summary_$N.txt
tail -n 5 $O/test_multi_$N.log
EXTRA="" run default A=1
EXTRA="--no_comm" run nocomm A=1
if [ "$MODE" = "mid" ]; then
  EXTRA="" run ov8_tail96 B200DDP_COMM_BLOCKS=8
  EXTRA="" run ov48_tail96 B200DDP_COMM_BLOCKS=48
  EXTRA="" run ov24_tail24 B200DDP_TAIL_BLOCKS=24
fi
if [ "$MODE" = "full" ]; then
  EXTRA="" run ov8_tail96 B200DDP_COMM_BLOCKS=8
  EXTRA="" run ov48_tail96 B200DDP_COMM_BLOCKS=48
  EXTRA="" run ov24_tail24 B200DDP_TAIL_BLOCKS=24
  EXTRA="" run ov24_tail128 B200DDP_TAIL_BLOCKS=128
  EXTRA="" run nonvls B200DDP_DISABLE_NVLS=1
  EXTRA="--wire_dtype fp32" run wirefp32 A=1
  EXTRA="--backend nccl --no_graph" run nccl_nograph A=1
  EXTRA="--no_graph" run nograph A=1
fi
if [ "$MODE" = "mid" ]; then cat $O/diag_summary_$N.txt; exit 0; fi
echo "== sweep"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 \
   bench/allreduce_sweep.py --max_mb 256 --out $O/sweep_$N.json > $O/sweep_$N.log 2>&1; echo "sweep rc=$?" | tee -a $O/diag_summary_$N.txt
tail -n 14 $O/sweep_$N.log
cat $O/diag_summary_$N.txt
