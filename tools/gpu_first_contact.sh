#!/bin/bash
# First-contact run on a 1-GPU box: each stage under its own timeout so a hung kernel cannot eat the lease.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvidia-smi.txt 2>&1
python - > $O/env.txt 2>&1 <<'PY'
import torch, sys
print(torch.__version__, torch.cuda.is_available(), torch.cuda.device_count())
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
sys.path.insert(0, ".")
from b200ddp import _ext
C = _ext.get(); print("ext", C.__file__)
PY
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > $O/test_kernels.log 2>&1; echo "kernels rc=$?" | tee -a $O/summary.txt
echo "== gemm";    timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q > $O/test_gemm.log 2>&1; echo "gemm rc=$?" | tee -a $O/summary.txt
echo "== bench ours"; timeout 900 python bench.py --gpus 1 --steps 30 --warmup 8 > $O/bench_ours.json 2> $O/bench_ours.err; echo "bench ours rc=$?" | tee -a $O/summary.txt
if ! grep -q '"value"' $O/bench_ours.json; then
  echo "== bench ours (tcgen05 linear disabled: fallback diagnosis)"
  B200DDP_DISABLE_TC=1 timeout 900 python bench.py --gpus 1 --steps 30 --warmup 8 > $O/bench_ours_notc.json 2> $O/bench_ours_notc.err; echo "bench ours notc rc=$?" | tee -a $O/summary.txt
  cat $O/bench_ours_notc.json; tail -n 5 $O/bench_ours_notc.err
fi
echo "== bench ours nograph"; timeout 600 python bench.py --gpus 1 --steps 30 --warmup 8 --no_graph --skip_e2e > $O/bench_ours_nograph.json 2> $O/bench_ours_nograph.err; echo "bench nograph rc=$?" | tee -a $O/summary.txt
echo "== bench ref"; timeout 900 python bench.py --impl reference --gpus 1 --steps 30 --warmup 8 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
tail -n 3 $O/smoke.log $O/test_kernels.log $O/test_gemm.log
cat $O/bench_ours.json $O/bench_ours_nograph.json $O/bench_ref.json
tail -n 5 $O/bench_ours.err $O/bench_ref.err
