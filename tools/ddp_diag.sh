#!/bin/bash
# N-GPU decomposition of the DDP step: tests, then the same step with / without communication, without the buffer broadcast,
# with different tail-bucket bounds, plus a CUPTI timeline per rank (tools/trace_digest.py).  Short timeouts everywhere.
#   gpurun --gpus N --timeout 900 -- bash tools/ddp_diag.sh N [tests|notests]
cd "$(dirname "$0")/.."
N=${1:-2}
TESTS=${2:-tests}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/ddpdiag_summary_$N.txt
run() {  # name ENV=.. (bench args in $EXTRA)
  local name=$1; shift
  timeout 150 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps ${STEPS:-60} --warmup 8 --skip_e2e $EXTRA > $O/ddpdiag_${name}_$N.json 2> $O/ddpdiag_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$O/ddpdiag_${name}_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],4),'ms', d['config'].get('transport'), 'mib', d['config'].get('ddp',{}).get('bucket_mib'), 'algos', d['config'].get('ddp',{}).get('algos'), 'blocks', d['config'].get('ddp',{}).get('blocks'))" 2>&1)" | tee -a $O/ddpdiag_summary_$N.txt
}
if [ "$TESTS" = "tests" ]; then
  echo "== tests"; B200DDP_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests (world $N) rc=$?" | tee -a $O/ddpdiag_summary_$N.txt
  tail -n 6 $O/test_multi_$N.log
fi
EXTRA="" run default A=1
EXTRA="" run default_again A=1
EXTRA="--no_comm" run nocomm A=1
EXTRA="--no_broadcast_buffers" run nobufbcast A=1
EXTRA="" run tail1mib B200DDP_TAIL_BUCKET_MB=1
EXTRA="" run tail2mib B200DDP_TAIL_BUCKET_MB=2
EXTRA="" run tail8mib B200DDP_TAIL_BUCKET_MB=8
EXTRA="" run tailblocks48 B200DDP_TAIL_BLOCKS=48
EXTRA="" run blocks12 B200DDP_COMM_BLOCKS=12
EXTRA="--trace_dir $O/trace_ddp_$N" STEPS=20 run traced A=1
python tools/trace_digest.py $O/trace_ddp_$N --label "$N GPUs, default" > $O/ddp_timeline_$N.md 2>&1
head -n 60 $O/ddp_timeline_$N.md
cat $O/ddpdiag_summary_$N.txt
