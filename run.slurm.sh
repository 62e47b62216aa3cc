#!/bin/bash
# Per-node half of run.sbatch: SLURM gives the node count and this node's index, torchrun spawns one rank per GPU.
set -eu
GPUS_PER_NODE=${GPUS_PER_NODE:-${SLURM_GPUS_ON_NODE:-8}}
exec python -m torch.distributed.run \
  --nnodes="${SLURM_JOB_NUM_NODES}" --node-rank="${SLURM_NODEID}" --nproc-per-node="${GPUS_PER_NODE}" \
  --master-addr "${MASTER_ADDR}" --master-port "${MASTER_PORT}" \
  ddp.py --backend "${B200DDP_BACKEND:-nccl}" "$@"
