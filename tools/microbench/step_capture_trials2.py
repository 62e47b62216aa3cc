#!/usr/bin/env python
"""What flips the captured ResNet-50 step into its slow mode (+0.17 ms = +0.4 us at every kernel boundary)?  Fresh captures
inside one process, no communication kernels at all, with memory statistics."""
import gc
import os
import sys

os.environ["B200DDP_DEBUG_BUCKET"] = "1"
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from b200ddp.engine.step import TrainStep
    from b200ddp.models import build_model
    from b200ddp.ops import MSELoss
    from b200ddp.optim import FusedSGD
    from b200ddp.parallel import DistributedDataParallel
    from b200ddp.utils import to_mixed_bf16
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = to_mixed_bf16(build_model("resnet50").to(dev)).to(memory_format=torch.channels_last)
    mode = os.environ.get("TRIAL_MODE", "ddp")
    if mode == "ddp":
        wrapped = DistributedDataParallel(model, device_ids=[local], backend="b200", broadcast_buffers=False)
    else:
        wrapped = model
    opt = FusedSGD(model.parameters(), lr=1e-3, max_grad_norm=1000.0)
    x = torch.randn(32, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.zeros(32, 1000, device=dev, dtype=torch.bfloat16)
    keep = []
    import subprocess, time
    pre = os.environ.get("TRIAL_PRE", "none")
    if pre == "idle":
        time.sleep(10.0)
    elif pre == "busy":
        a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        t_end = time.time() + 10.0
        while time.time() < t_end:
            for _ in range(20):
                a @ a
            torch.cuda.synchronize()
    t_start = time.time()
    for trial in range(int(os.environ.get("TRIALS", "7"))):
        if trial == 5:
            gc.collect(); torch.cuda.empty_cache()
        step = TrainStep(wrapped, MSELoss(), opt, dev, use_graph=True, graph_warmup=1)
        for _ in range(4):
            step(x, y)
        torch.cuda.synchronize(); dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step(x, y)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,clocks.gr,clocks.video,pstate,power.draw,temperature.gpu", "--format=csv,noheader", "-i", str(local)],
                                 capture_output=True, text=True).stdout.strip()
            print(f"[{time.time() - t_start:5.1f}s] {smi} | ", end="")
            print(f"mode={mode} trial {trial}: {float(t):.4f} ms  reserved {torch.cuda.memory_reserved(dev) / 2**30:.2f} GiB allocated {torch.cuda.memory_allocated(dev) / 2**30:.2f} GiB "
                  f"static_x ptr {step._static_x.data_ptr():#x}", flush=True)
        if os.environ.get("KEEP_STEPS") == "1":
            keep.append(step)
        del step
    if keep:
        # is fast / slow a property of the graph instance?  re-time every kept graph, twice, in order
        for rnd in range(2):
            out = []
            for st in keep:
                for _ in range(2):
                    st(x, y)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30):
                    st(x, y)
                e1.record()
                torch.cuda.synchronize()
                out.append(round(e0.elapsed_time(e1) / 30, 4))
            if rank == 0:
                print(f"re-timed kept graphs, round {rnd}: {out}", flush=True)
    from b200ddp.parallel.peer import PeerCollectives
    PeerCollectives.shutdown_all()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
