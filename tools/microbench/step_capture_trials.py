#!/usr/bin/env python
"""Does ONE extra kernel node in the captured ResNet-50 step change the step time, and is the effect tied to the node's
parameter size or random per capture?  Several fresh captures per variant inside one process (bucket kernels skipped)."""
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
    ddp = DistributedDataParallel(model, device_ids=[local], backend="b200", broadcast_buffers=True)
    C = ddp.comm.C
    opt = FusedSGD(model.parameters(), lr=1e-3, max_grad_norm=1000.0)
    x = torch.randn(32, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.zeros(32, 1000, device=dev, dtype=torch.bfloat16)
    small = torch.ones(64, device=dev)
    bufs = [b.data for b in model.buffers()]
    variants = {
        "no extra node": lambda: None,
        "tiny-parameter peer kernel (barrier, 56 B of arguments)": lambda: C.peer_barrier(ddp.comm.arena, 1, 1, 30.0),
        "3 KB-parameter peer kernel (broadcast of 1 tensor)": lambda: ddp.comm.broadcast_tensors([small], src=0),
        "3 KB-parameter peer kernel (broadcast of 159 buffers)": lambda: ddp.comm.broadcast_tensors(bufs, src=0),
        "plain ATen kernel": lambda: small.add_(1.0),
    }
    for name, op in variants.items():
        ddp._sync_buffers_after_forward = op
        times = []
        for trial in range(4):
            step = TrainStep(ddp, MSELoss(), opt, dev, use_graph=True, graph_warmup=1)
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
            times.append(round(float(t), 4))
            del step
        if rank == 0:
            print(f"{name:60s} {times}", flush=True)
    from b200ddp.parallel.peer import PeerCollectives
    PeerCollectives.shutdown_all()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
