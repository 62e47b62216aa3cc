#!/usr/bin/env python
"""Why does every kernel boundary of a captured step cost ~0.5 us more once the graph contains a gradient-exchange kernel?
Times the replay of a chain of 400 tiny kernels captured (a) alone, (b) with ONE extra node of various kinds in the middle.

  torchrun --nproc-per-node 2 tools/microbench/graph_gap_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from b200ddp.parallel.peer import PeerCollectives
    comm = PeerCollectives.get(None, dev, min_bytes=64 << 20)
    C = comm.C
    x = torch.zeros(int(os.environ.get('PROBE_ELEMS', str(1 << 24))), device=dev)    # 64 MB: ~25 us per kernel, so launch-ahead matters
    sym = comm.symmetric_empty(1 << 16, torch.float32)
    sym.zero_()
    pinned = torch.zeros(1024).pin_memory()
    red = torch.ones(4096, device=dev)
    N = 200

    def chain(extra):
        for i in range(N):
            x.add_(1.0)
            if i == N // 2 and extra is not None:
                extra()

    variants = {
        "plain chain": None,
        "+ kernel on arena (VMM, peer-mapped, multicast-bound) memory": lambda: sym.add_(1.0),
        "+ peer barrier kernel (st.release.sys / ld.acquire.sys on peers)": lambda: C.peer_barrier(comm.arena, 1, 1, 30.0),
        "+ fused bucket allreduce kernel (two_shot: P2P ld/st.sys)": lambda: comm.allreduce_([red], wire="fp32", algo="two_shot", scale=0.5),
        "+ fused bucket allreduce kernel (nvls: multimem)": (lambda: comm.allreduce_([red], wire="fp32", algo="nvls", scale=0.5)) if comm.nvls else None,
        "+ peer broadcast kernel": lambda: comm.broadcast_tensors([red], src=0),
        "+ memcpy to pinned host memory": lambda: pinned.copy_(red[:1024], non_blocking=True),
        "+ 2nd stream fork/join around one tiny kernel": "fork",
    }
    side = torch.cuda.Stream(device=dev)
    for name, extra in variants.items():
        if extra == "fork":
            def extra():
                ev = torch.cuda.Event(); ev.record()
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    y = red * 1.0
                    ev2 = torch.cuda.Event(); ev2.record()
                torch.cuda.current_stream().wait_event(ev2)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            chain(extra)                                   # warm-up
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain(extra)
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        per = e0.elapsed_time(e1) * 1e3 / 10 / N
        if rank == 0:
            print(f"{name:70s} {per:6.3f} us per kernel", flush=True)
        del g
    PeerCollectives.shutdown_all()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
