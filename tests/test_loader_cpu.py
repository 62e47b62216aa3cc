"""BatchLoader: ordered hand-out with several helper threads, tail batch, early close, error propagation."""
import pytest
import torch

from b200ddp.data import BatchLoader, FooDataset
from b200ddp.parallel import ShardedSampler


def _threaded(ds, **kw):
    loader = BatchLoader(ds, pin_memory=False, **kw)
    loader.background = True          # the helper-thread path is normally tied to pinned (CUDA) runs
    return loader


def test_threaded_loader_preserves_order_and_tail():
    torch.manual_seed(0)
    ds = FooDataset(1000)
    for workers in (1, 3):
        out = list(_threaded(ds, batch_size=32, workers=workers))
        assert len(out) == 32 and out[-1][0].shape[0] == 1000 - 31 * 32
        assert torch.equal(torch.cat([b[0] for b in out]), ds.X)
        assert torch.equal(torch.cat([b[1] for b in out]), ds.Y)
    assert len(list(_threaded(ds, batch_size=32, drop_last=True))) == 31


def test_threaded_loader_follows_sampler_and_closes_early():
    ds = FooDataset(257)
    sampler = ShardedSampler(ds, num_replicas=2, rank=1, seed=5)
    sampler.set_epoch(3)
    expect = list(sampler)
    loader = _threaded(ds, batch_size=16, sampler=sampler)
    it = iter(loader)
    first = next(it)
    assert torch.equal(first[0], ds.X[torch.tensor(expect[:16])])
    it.close()                        # must stop the helper threads without hanging
    got = torch.cat([b[0] for b in loader])
    assert torch.equal(got, ds.X[torch.tensor(expect)])


def test_threaded_loader_surfaces_errors():
    class Broken(FooDataset):
        def batch(self, idx):
            raise ValueError("boom")

    with pytest.raises(ValueError, match="boom"):
        list(_threaded(Broken(64), batch_size=8))


def test_endless_sampler_streams_across_epochs():
    from b200ddp.parallel import EndlessSampler
    ds = FooDataset(40)
    base = ShardedSampler(ds, num_replicas=2, rank=0, seed=1)
    loader = _threaded(ds, batch_size=8, sampler=EndlessSampler(base), drop_last=True)
    expect = []
    for epoch in range(3):
        base.set_epoch(epoch)
        expect += list(base)
    it = iter(loader)
    rows = torch.cat([next(it)[0] for _ in range(7)])     # 20 indices per epoch -> crosses two epoch boundaries
    it.close()
    assert torch.equal(rows, ds.X[torch.tensor(expect[:56])])
