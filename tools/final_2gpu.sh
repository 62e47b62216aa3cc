#!/bin/bash
# 2-GPU acceptance: transport + DDP tests at world 2, then one CUPTI-traced run of the default configuration for the
# per-rank timeline (profiles/ddp_timeline_r2.md).  A traced run is never a timing source.
cd "$(dirname "$0")/.."
N=2; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_data_parallel.py -m gpu -q > $O/test_multi_2.log 2>&1; echo "multi tests (world 2) rc=$? : $(tail -n 1 $O/test_multi_2.log)"
grep -n "AssertionError\|^E  " $O/test_multi_2.log | head -8
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --warmup 8 --skip_e2e"
rm -rf $O/trace_ddp_r2_$N $O/trace_ddp_r2_nocomm_$N
timeout 200 $R --steps 20 --trace_dir $O/trace_ddp_r2_$N > $O/traced_r2.json 2> $O/traced_r2.err; echo "traced rc=$?"
timeout 200 $R --steps 20 --no_comm --trace_dir $O/trace_ddp_r2_nocomm_$N > $O/traced_r2_nocomm.json 2> $O/traced_r2_nocomm.err; echo "traced nocomm rc=$?"
python tools/trace_digest.py $O/trace_ddp_r2_$N --label "$N GPUs, round-2 default" > $O/ddp_timeline_r2_$N.md 2>&1
python tools/trace_diff.py $O/trace_ddp_r2_$N/rank0.json $O/trace_ddp_r2_nocomm_$N/rank0.json > $O/trace_diff_r2_$N.txt 2>&1
head -12 $O/trace_diff_r2_$N.txt
grep -n "exposed tail\|bucket_allreduce\|peer_broadcast" $O/ddp_timeline_r2_$N.md | head -20
