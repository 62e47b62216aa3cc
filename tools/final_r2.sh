#!/bin/bash
# Round-2 closing run on 1 GPU: GPU test tier, smoke(), headline bench (ours, reference, stock eager / graphed), a CUPTI launch list of
# one step, and ncu --set full captures of the stem kernels.  Short timeouts everywhere; results in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
echo "== gpu tests"; timeout 500 python -m pytest tests -m gpu -q > $O/test_gpu_all.log 2>&1; echo "gpu tests rc=$? : $(tail -n 1 $O/test_gpu_all.log)"
grep -n "^FAILED\|^ERROR" $O/test_gpu_all.log | head -10
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$? : $(tail -n 1 $O/smoke.log | cut -c1-200)"
show() { python -c "import json; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', round(d.get('value',0),1), d.get('unit'), round(d.get('ms_per_step',0),4),'ms  e2e', round((d.get('e2e') or {}).get('value',0),1), 'launches', d.get('gpu_launches'), d.get('clocks'))" 2>&1; }
timeout 300 python bench.py --gpus 1 --steps 60 --warmup 8 > $O/bench_ours.json 2> $O/bench_ours.err; echo "rc=$?"; show $O/bench_ours.json ours
echo "== ncu stem kernels"
timeout 200 $NCU --set full --import-source on -k regex:conv_tap_gemm_kernel -s 2 -c 1 -o $O/prof_stem_fprop -f python bench/stem_bench.py --reps 2 > $O/prof_stem_fprop.log 2>&1; echo "ncu stem fprop rc=$?"
ls -la $O/prof_stem_*.ncu-rep 2>&1 | cut -c1-120
echo "== launch list (CUPTI trace of the captured step; not a timing source)"
rm -rf $O/trace_final_1; timeout 200 python bench.py --gpus 1 --steps 10 --warmup 8 --skip_e2e --trace_dir $O/trace_final_1 > $O/traced_final.json 2> $O/traced_final.err; echo "trace rc=$?"
python tools/trace_kernels.py $O/trace_final_1/rank0.json > $O/launches_trace_r2.txt 2>&1; head -n 14 $O/launches_trace_r2.txt | cut -c1-150
