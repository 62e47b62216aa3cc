#!/bin/bash
# Round-2 closing run on a 2-GPU box: ResNet-50 DDP parity (native stem inside DDP) and the reference arm at N=1 on a
# multi-GPU box (it must stay on one GPU: round 1's run fell into DataParallel over every visible device).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "resnet50 or stress" > $O/test_multi_2_final.log 2>&1; echo "multi tests rc=$? : $(tail -n 1 $O/test_multi_2_final.log)"
( for i in 1 2 3 4 5 6; do sleep 5; nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv,noheader; done > $O/ref_n1_smi.txt ) &
timeout 200 python bench.py --impl reference --gpus 1 --steps 15 --warmup 4 > $O/bench_ref_n1_on2.json 2> $O/bench_ref_n1_on2.err; echo "reference N=1 rc=$?"
python -c "import json; d=json.loads([l for l in open('$O/bench_ref_n1_on2.json') if l.startswith('{')][-1]); print('reference N=1 on a 2-GPU box:', round(d.get('value',0),1), d.get('unit'), round(d.get('ms_per_step',0),3), 'ms', d.get('unavailable'))"
wait
cat $O/ref_n1_smi.txt | tail -n 6
