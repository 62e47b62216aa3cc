"""Process-group helpers (parity: reference ``utils.py:84-101``) plus the
environment-driven setup the reference does inline in ``ddp.py:80-115``."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _pg_ready() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if _pg_ready() else 0


def get_world_size() -> int:
    return dist.get_world_size() if _pg_ready() else 1


def is_main_process() -> bool:
    return get_rank() == 0


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ[name])
    except (KeyError, ValueError):
        return default


def resolve_local_rank(cli_value: int = -1) -> int:
    """env ``LOCAL_RANK`` wins over ``--local_rank`` (reference ``ddp.py:85``)."""
    return env_int("LOCAL_RANK", cli_value)


def barrier_all() -> None:
    if _pg_ready() and dist.get_world_size() > 1:
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """All-reduce(MAX) of a host scalar; multi-GPU timings are max over ranks."""
    if not _pg_ready() or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
