#!/bin/bash
# N-GPU evidence run (default 8): transport + DDP tests at world N, headline bench (ours / no-comm / stock reference),
# allreduce sweep 1 KB .. 1 GB.  Short timeouts everywhere; every result lands in gpurun_out/.
cd "$(dirname "$0")/.."
N=${1:-8}
SWEEP_MB=${2:-1024}
O=gpurun_out; mkdir -p $O
S=$O/scale_summary_$N.txt; rm -f $S
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
echo "== tests world=$N"
B200DDP_TEST_WORLD=$N timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/test_multi_$N.log 2>&1; echo "multi tests (world $N) rc=$? : $(tail -n 1 $O/test_multi_$N.log)" | tee -a $S
run() {  # name, extra args...
  local name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
     bench.py --gpus $N --steps 60 --warmup 8 "$@" > $O/bench_${name}_$N.json 2> $O/bench_${name}_$N.err
  echo "bench $name rc=$? $(python -c "import json; d=json.loads([l for l in open('$O/bench_${name}_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],4),'ms  e2e', round((d.get('e2e') or {}).get('value',0),1), d['config'].get('transport'), d['config'].get('ddp',{}).get('algos'), d['config'].get('ddp',{}).get('blocks'), d.get('clocks'))" 2>&1)" | tee -a $S
}
run ours
run nocomm --no_comm --skip_e2e
run reference --impl reference
echo "== sweep"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 \
   bench/allreduce_sweep.py --max_mb $SWEEP_MB --blocks 32,296 --out $O/sweep_$N.json > $O/sweep_$N.log 2>&1; echo "sweep rc=$?" | tee -a $S
tail -n 30 $O/sweep_$N.log | cut -c1-230
cat $S
tail -n 3 $O/bench_ours_$N.err
