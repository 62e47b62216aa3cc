#!/bin/bash
# Stock arm (this repo's model definitions are NOT used: torchvision-shaped ResNet-50 + torch DDP/NCCL + ATen kernels), eager and CUDA-graphed.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
show() { python -c "import json; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', round(d.get('value',0),1), d.get('unit'), round(d.get('ms_per_step',0),4),'ms', d.get('unavailable'))" 2>&1; }
timeout 100 python bench.py --impl stock --gpus 1 --steps 30 --warmup 6 --skip_e2e > $O/bench_stock_eager.json 2> $O/bench_stock_eager.err; echo "rc=$?"; show $O/bench_stock_eager.json stock_eager
timeout 100 python bench.py --impl stock --stock_graph --gpus 1 --steps 30 --warmup 6 --skip_e2e > $O/bench_stock_graph.json 2> $O/bench_stock_graph.err; echo "rc=$?"; show $O/bench_stock_graph.json stock_graph
tail -n 2 $O/bench_stock_graph.err | cut -c1-200
