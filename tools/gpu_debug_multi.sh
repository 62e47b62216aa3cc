#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
export B200DDP_TEST_WORLD=$N
export B200DDP_TEST_DUMP_AFTER=45
timeout 110 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k test_peer_collectives > $O/dbg2_collectives_$N.log 2>&1; echo "rc=$?"
grep -vE "^$" $O/dbg2_collectives_$N.log | tail -n 80 | cut -c1-200
