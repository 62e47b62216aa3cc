cd /root/repo; O=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench/allreduce_sweep.py --max_mb 1024 --iters 10 --blocks 8,296 --out $O/sweep_8.json > $O/sweep_8.log 2>&1; echo "sweep rc=$?"
grep " B  nccl" $O/sweep_8.log | cut -c1-240
B200DDP_TEST_WORLD=8 timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k resnet50 > $O/test_multi_8_resnet.log 2>&1; echo "resnet50 parity (world 8) rc=$? : $(tail -n 1 $O/test_multi_8_resnet.log)"
grep -n "AssertionError" $O/test_multi_8_resnet.log | head -3
