#!/bin/bash
# N-GPU acceptance run: transport tests at world=N, headline bench (both arms), allreduce sweep.  Short timeouts everywhere.
cd "$(dirname "$0")/.."
N=${1:-8}
SWEEP_MB=${2:-64}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/scale_summary_$N.txt
echo "== tests world=$N"; timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests rc=$?" | tee -a $O/scale_summary_$N.txt
tail -n 4 $O/test_multi_$N.log
for impl in ours reference; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
     bench.py --impl $impl --gpus $N --steps 30 --warmup 8 > $O/bench_${impl}_$N.json 2> $O/bench_${impl}_$N.err
  echo "bench $impl rc=$? $(python -c "import json; d=json.loads([l for l in open('$O/bench_${impl}_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],3),'ms  e2e', round(d.get('e2e',{}).get('value',0),1), d['config'].get('transport'), d['config'].get('ddp',{}).get('algos'))" 2>&1)" | tee -a $O/scale_summary_$N.txt
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 \
   bench.py --gpus $N --steps 30 --warmup 8 --skip_e2e --no_comm > $O/bench_nocomm_$N.json 2> $O/bench_nocomm_$N.err
echo "bench nocomm rc=$? $(python -c "import json; d=json.loads([l for l in open('$O/bench_nocomm_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],3),'ms')" 2>&1)" | tee -a $O/scale_summary_$N.txt
echo "== sweep"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 \
   bench/allreduce_sweep.py --max_mb $SWEEP_MB --out $O/sweep_$N.json > $O/sweep_$N.log 2>&1; echo "sweep rc=$?" | tee -a $O/scale_summary_$N.txt
tail -n 12 $O/sweep_$N.log | cut -c1-260
cat $O/scale_summary_$N.txt
tail -n 5 $O/bench_ours_$N.err
