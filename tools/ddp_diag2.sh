#!/bin/bash
# N-GPU: in-line (wide) vs overlapped (light) gradient exchange under the CUDA-graph step.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/ddpdiag2_summary_$N.txt
run() {
  local name=$1; shift
  timeout 150 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps ${STEPS:-60} --warmup 8 --skip_e2e $EXTRA > $O/ddpdiag2_${name}_$N.json 2> $O/ddpdiag2_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$O/ddpdiag2_${name}_$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'samples/s', round(d['ms_per_step'],4),'ms', d['config'].get('transport'), 'mib', d['config'].get('ddp',{}).get('bucket_mib'))" 2>&1)" | tee -a $O/ddpdiag2_summary_$N.txt
}
if [ "${2:-tests}" = "tests" ]; then
  B200DDP_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/test_multi_$N.log 2>&1; echo "multi tests (world $N) rc=$?" | tee -a $O/ddpdiag2_summary_$N.txt
  tail -n 4 $O/test_multi_$N.log
fi
EXTRA="--no_comm" run nocomm A=1
EXTRA="" run inline296 A=1
EXTRA="" run inline148 B200DDP_WIDE_BLOCKS=148
EXTRA="" run inline296_cap100 B200DDP_BUCKET_CAP_MB=100
EXTRA="--bucket_cap_mb 100" run inline296_capflag A=1
EXTRA="" run overlap B200DDP_DDP_SERIAL=0
EXTRA="--no_comm" run nocomm_again A=1
EXTRA="--trace_dir $O/trace_ddp_inline_$N" STEPS=20 run traced A=1
python tools/trace_digest.py $O/trace_ddp_inline_$N --label "$N GPUs, in-line wide bucket kernels" > $O/ddp_timeline_inline_$N.md 2>&1
head -n 22 $O/ddp_timeline_inline_$N.md
cat $O/ddpdiag2_summary_$N.txt
