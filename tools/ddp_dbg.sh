#!/bin/bash
# 2-GPU A/B of the gradient-exchange knobs (no-comm vs default vs tail algorithm / tail bound); results quoted in profiles/ddp_overhead_r2.txt.
cd /root/repo
N=${1:-2}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 60 --warmup 8 --skip_e2e"
show() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', round(d['ms_per_step'],4), d['config'].get('ddp',{}).get('bucket_mib'), d['config'].get('ddp',{}).get('algos'), d['config'].get('ddp',{}).get('blocks'))"; }
timeout 150 $R --no_comm > gpurun_out/dbg_nocomm.json 2>/dev/null; show gpurun_out/dbg_nocomm.json nocomm
timeout 150 $R > gpurun_out/dbg_def.json 2>/dev/null; show gpurun_out/dbg_def.json default_tail_one_shot
B200DDP_TAIL_ONE_SHOT_MAX_MB=0 timeout 150 $R > gpurun_out/dbg_two.json 2>/dev/null; show gpurun_out/dbg_two.json tail_two_shot
B200DDP_TAIL_BUCKET_MB=2 timeout 150 $R > gpurun_out/dbg_t2.json 2>/dev/null; show gpurun_out/dbg_t2.json tail2mib_one_shot
B200DDP_TAIL_BUCKET_MB=8 timeout 150 $R > gpurun_out/dbg_t8.json 2>/dev/null; show gpurun_out/dbg_t8.json tail8mib_one_shot
timeout 150 $R > gpurun_out/dbg_def2.json 2>/dev/null; show gpurun_out/dbg_def2.json default_again
