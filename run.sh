#!/bin/sh
# Single-node launch, one process per GPU (reference: run.sh, 4 GPUs via torch.distributed.launch on port 9315).
# torchrun is used because the legacy launcher's --local-rank spelling breaks the reference on torch >= 2;
# ddp.py here accepts both spellings and reads LOCAL_RANK from the environment first.
set -eu
NGPU=${NGPU:-$(nvidia-smi -L 2>/dev/null | wc -l)}
[ "$NGPU" -ge 1 ] || NGPU=1
PORT=${MASTER_PORT:-9315}
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node="$NGPU" \
     --master-addr 127.0.0.1 --master-port "$PORT" ddp.py "$@"
