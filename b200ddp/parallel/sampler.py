"""Sharded sampler.  Parity target: ``torch.utils.data.DistributedSampler`` as
used by the reference (``ddp.py:138-141``, ``set_epoch`` at ``ddp.py:214``):
epoch-seeded permutation, pad to a multiple of the world size by wrapping,
then every ``world``-th index starting at ``rank``.  Index sets are identical
to torch's for the same (seed, epoch) so loss curves can be compared 1:1.

Extra over the reference: ``start_index`` for mid-epoch resume.
"""
from __future__ import annotations

import math
from typing import Iterator, Optional, Sized

import torch
import torch.distributed as dist


class ShardedSampler(torch.utils.data.Sampler):
    def __init__(self, dataset: Sized, num_replicas: Optional[int] = None, rank: Optional[int] = None,
                 shuffle: bool = True, seed: int = 0, drop_last: bool = False):
        if num_replicas is None:
            num_replicas = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if not 0 <= rank < num_replicas:
            raise ValueError(f"rank {rank} outside [0, {num_replicas})")
        self.n = len(dataset)
        self.world = num_replicas
        self.rank = rank
        self.shuffle = shuffle
        self.seed = seed
        self.drop_last = drop_last
        self.epoch = 0
        self.start_index = 0
        if drop_last and self.n % self.world:
            self.per_rank = math.ceil((self.n - self.world) / self.world)
        else:
            self.per_rank = math.ceil(self.n / self.world)
        self.padded = self.per_rank * self.world

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def set_start_index(self, start: int) -> None:
        """Skip the first ``start`` samples of this rank's shard (resume)."""
        self.start_index = int(start)

    def global_order(self) -> torch.Tensor:
        if self.shuffle:
            gen = torch.Generator()
            gen.manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=gen)
        else:
            order = torch.arange(self.n)
        if self.drop_last:
            return order[: self.padded]
        short = self.padded - order.numel()
        if short > 0:
            reps = math.ceil(short / max(1, order.numel()))
            order = torch.cat([order, order.repeat(reps)[:short]])
        return order

    def shard(self) -> torch.Tensor:
        return self.global_order()[self.rank:self.padded:self.world]

    def __iter__(self) -> Iterator[int]:
        mine = self.shard()
        start, self.start_index = self.start_index, 0
        return iter(mine[start:].tolist())

    def __len__(self) -> int:
        return self.per_rank


class EndlessSampler(torch.utils.data.Sampler):
    """Chains epochs of a finite sampler into one endless index stream (``set_epoch`` is advanced on every pass), so a
    loader's helper threads and prefetch queue never drain at an epoch boundary.  Meant for throughput runs over small
    synthetic datasets where an "epoch" is only a handful of batches."""

    def __init__(self, base, start_epoch: int = 0):
        self.base = base
        self.epoch = start_epoch

    def __iter__(self):
        while True:
            if hasattr(self.base, "set_epoch"):
                self.base.set_epoch(self.epoch)
            n = 0
            for idx in self.base:
                n += 1
                yield idx
            if n == 0:
                return
            self.epoch += 1

    def __len__(self) -> int:
        return 1 << 62
