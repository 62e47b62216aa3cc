"""Multi-GPU checks of the NVSwitch peer-memory transport (run with >= 2 GPUs)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _world():
    """World size for these tests: 2 by default (every box with >= 2 GPUs); B200DDP_TEST_WORLD=4|8 widens it."""
    want = int(os.environ.get("B200DDP_TEST_WORLD", "2"))
    return max(2, min(want, torch.cuda.device_count()))


def _entry(rank, fn, world, port, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), B200DDP_TIMEOUT_S="6")
    torch.cuda.set_device(rank)
    if os.environ.get("B200DDP_TEST_DUMP_AFTER"):          # debugging aid: dump every rank's Python stack if a test stalls
        import faulthandler
        import sys
        faulthandler.dump_traceback_later(float(os.environ["B200DDP_TEST_DUMP_AFTER"]), exit=True, file=sys.stderr)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
        torch.cuda.synchronize()
    except BaseException:
        # a failing rank must not wait for its peers in collective teardown (they may be blocked on it): report and leave
        import traceback
        traceback.print_exc()
        os._exit(1)
    try:
        from b200ddp.parallel.peer import PeerCollectives
        PeerCollectives.shutdown_all()
    finally:
        dist.destroy_process_group()


def _spawn(fn, port, *args, world=None):
    mp.spawn(_entry, args=(fn, world or _world(), port, args), nprocs=world or _world(), join=True)


# --------------------------------------------------------------------------------------------------
def _check_collectives(rank, world):
    from b200ddp.parallel.peer import PeerCollectives
    dev = torch.device("cuda", rank)
    comm = PeerCollectives.get(None, dev, min_bytes=64 << 20)
    algos = ["one_shot", "two_shot"] + (["nvls", "nvls_one_shot"] if comm.nvls else [])
    print(f"[rank {rank}] nvls={comm.nvls} world={world}", flush=True)

    def stage(msg):
        if rank == 0:
            print(f"[collectives] {msg}", flush=True)
    torch.manual_seed(100 + rank)
    sizes = [1, 7, 165, 4096, 100003, 1 << 20]
    for dtype in (torch.float32, torch.bfloat16):
        for wire in ("fp32", "bf16"):
            for algo in algos:
                if algo in ("one_shot", "nvls_one_shot"):
                    use = sizes[:4]
                else:
                    use = sizes
                stage(f"allreduce {dtype} wire={wire} algo={algo}")
                tensors = [torch.randn(n, device=dev).to(dtype) for n in use]
                ref = [t.float().clone() for t in tensors]
                for r in ref:
                    dist.all_reduce(r)
                comm.allreduce_(tensors, wire=wire, algo=algo, scale=1.0 / world)
                torch.cuda.synchronize()
                comm.check()
                lossy = wire == "bf16" or dtype == torch.bfloat16
                for t, r in zip(tensors, ref):
                    exp = r / world
                    tol = 3e-2 if lossy else 1e-5
                    err = (t.float() - exp).abs().max().item()
                    assert err <= tol * max(1.0, exp.abs().max().item()), (dtype, wire, algo, t.numel(), err)
                # every rank must hold bit-identical results
                flat = torch.cat([t.float().reshape(-1) for t in tensors])
                gathered = [torch.empty_like(flat) for _ in range(world)]
                dist.all_gather(gathered, flat)
                assert all(torch.equal(g, gathered[0]) for g in gathered), (dtype, wire, algo)
    stage("symmetric")
    # in-place allreduce on symmetric (arena-resident) tensors
    for dtype in (torch.float32, torch.bfloat16):
        sym = comm.symmetric_empty(40000, dtype)
        for algo in ["two_shot"] + (["nvls"] if comm.nvls else []):
            torch.manual_seed(50 + rank)
            sym.copy_(torch.randn(40000, device=dev).to(dtype))
            ref = sym.float().clone()
            dist.all_reduce(ref)
            dist.barrier(device_ids=[rank])
            comm.allreduce_symmetric_(sym, algo=algo, scale=1.0 / world)
            torch.cuda.synchronize()
            comm.check()
            tol = 1e-5 if dtype == torch.float32 else 3e-2
            assert (sym.float() - ref / world).abs().max().item() <= tol * max(1.0, ref.abs().max().item() / world), (dtype, algo)
    stage("broadcast")
    # broadcast: odd sizes, unaligned views, several dtypes, a tensor larger than one staging chunk
    torch.manual_seed(7)
    base = [torch.randn(5), torch.randn(1000, 33), torch.randint(0, 100, (77,)), torch.randn(3, 5, 7).to(torch.bfloat16),
            torch.randn(20 * 1024 * 1024)]
    mine = [(b.clone() if rank == 0 else torch.zeros_like(b)).to(dev) for b in base]
    view = torch.zeros(1001, device=dev)[1:]                 # 4-byte aligned only
    if rank == 0:
        view.copy_(torch.arange(1000.0))
    os.environ["B200DDP_SCRATCH_MB"] = "64"
    n = comm.broadcast_tensors(mine + [view], src=0)
    torch.cuda.synchronize()
    comm.check()
    assert n >= 2
    for m, b in zip(mine, base):
        assert torch.equal(m.cpu(), b), (rank, b.shape)
    assert torch.equal(view.cpu(), torch.arange(1000.0))
    stage("stress")
    # stress the barrier protocol: many back-to-back tiny collectives
    t = torch.ones(3, device=dev)
    for _ in range(200):
        comm.allreduce_([t], wire="fp32", algo="one_shot", scale=1.0 / world)
    torch.cuda.synchronize()
    comm.check()
    assert torch.allclose(t, torch.ones_like(t))


def test_peer_collectives(free_port):
    _spawn(_check_collectives, free_port)


# --------------------------------------------------------------------------------------------------
def _train_pair(rank, world, model_fn, make_batch, steps, ddp_kwargs, dtype, graph):
    """Train the same model with our DDP (b200 transport) and with stock torch DDP (NCCL); compare."""
    from b200ddp.engine.step import TrainStep
    from b200ddp.ops import MSELoss
    from b200ddp.optim import FusedSGD
    from b200ddp.parallel import DistributedDataParallel
    dev = torch.device("cuda", rank)
    torch.manual_seed(1000 + rank)                           # rank-dependent init: wrap must broadcast rank 0's
    ours = model_fn().to(dev)
    if dtype == torch.bfloat16:
        ours = ours.to(dtype)
    torch.manual_seed(1000)                                  # rank 0's init everywhere for the stock model
    stock = model_fn().to(dev)
    if dtype == torch.bfloat16:
        stock = stock.to(dtype)
    ddp = DistributedDataParallel(ours, device_ids=[rank], backend="b200", **ddp_kwargs)
    assert ddp.backend_name == "b200"
    for a, b in zip(ours.parameters(), stock.parameters()):
        assert torch.equal(a, b), "peer broadcast did not deliver rank 0's weights"
    ref = nn.parallel.DistributedDataParallel(stock, device_ids=[rank])
    opt = FusedSGD(ours.parameters(), lr=0.05, max_grad_norm=1.0)
    ropt = torch.optim.SGD(stock.parameters(), lr=0.05)
    step = TrainStep(ddp, MSELoss(), opt, dev, use_graph=graph)
    for i in range(steps):
        x, y = make_batch(rank, i, dev, dtype)
        step(x, y)
        ropt.zero_grad(set_to_none=True)
        nn.functional.mse_loss(ref(x).float(), y.float()).backward()
        torch.nn.utils.clip_grad_norm_(stock.parameters(), 1.0)
        ropt.step()
    torch.cuda.synchronize()
    ddp.comm.check()
    tol = 1e-5 if dtype == torch.float32 else 5e-2
    for a, b in zip(ours.parameters(), stock.parameters()):
        assert torch.allclose(a.float(), b.float(), atol=tol, rtol=tol), (a.float() - b.float()).abs().max()
    # ranks agree bit-for-bit with each other
    flat = torch.cat([p.detach().float().reshape(-1) for p in ours.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    return ddp, step


def _foo_batch(rank, i, dev, dtype):
    g = torch.Generator().manual_seed(rank * 1000 + i)
    return torch.randn(32, 10, generator=g).to(dev, dtype), torch.randn(32, 5, generator=g).to(dev, dtype)


def _check_ddp_foo(rank, world):
    from b200ddp.models import FooModel
    for graph in (False, True):
        for kwargs in (dict(find_unused_parameters=True), dict(gradient_as_bucket_view=True),
                       dict(wire_dtype="fp32", bucket_cap_mb=0.0001, first_bucket_mb=0.00005)):
            ddp, step = _train_pair(rank, world, FooModel, _foo_batch, 12, kwargs, torch.float32, graph)
            stats = ddp.ddp_stats()
            assert stats["buckets_launched"] >= len(ddp._specs)
            if graph:
                assert step.graph is not None


def test_ddp_foo_matches_stock_ddp(free_port):
    _spawn(_check_ddp_foo, free_port)


def _check_ddp_mlp_bf16_and_unused(rank, world):
    from b200ddp.models import BranchyFooModel
    from b200ddp.ops import Linear
    from b200ddp.parallel import DistributedDataParallel

    def mlp():
        return nn.Sequential(Linear(256, 512, activation="relu"), Linear(512, 512, activation="relu"), Linear(512, 64))

    def batch(rank, i, dev, dtype):
        g = torch.Generator().manual_seed(rank * 77 + i)
        return torch.randn(64, 256, generator=g).to(dev, dtype), torch.randn(64, 64, generator=g).to(dev, dtype)

    _train_pair(rank, world, mlp, batch, 6, dict(bucket_cap_mb=0.5), torch.bfloat16, False)

    # unused-parameter semantics on the native reducer
    dev = torch.device("cuda", rank)
    torch.manual_seed(3)
    m = BranchyFooModel().to(dev)
    ddp = DistributedDataParallel(m, device_ids=[rank], backend="b200", find_unused_parameters=True)
    x, y = torch.randn(8, 10, device=dev), torch.randn(8, 5, device=dev)
    nn.functional.mse_loss(ddp(x, use_aux=False), y).backward()
    torch.cuda.synchronize()
    assert m.aux.weight.grad is None and m.net1.weight.grad is not None
    m.zero_grad(set_to_none=True)
    nn.functional.mse_loss(ddp(x, use_aux=(rank == 0)), y).backward()
    torch.cuda.synchronize()
    g = m.aux.weight.grad
    assert g is not None and g.abs().sum() > 0
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    assert all(torch.equal(t, gathered[0]) for t in gathered)
    # strict mode raises
    ddp2 = DistributedDataParallel(BranchyFooModel().to(dev), device_ids=[rank], backend="b200", find_unused_parameters=False)
    with pytest.raises(RuntimeError, match="find_unused_parameters"):
        nn.functional.mse_loss(ddp2(x, use_aux=False), y).backward()


def test_ddp_bf16_and_unused_parameters(free_port):
    _spawn(_check_ddp_mlp_bf16_and_unused, free_port)


# --------------------------------------------------------------------------------------------------
def _check_timeout(rank, world):
    """A peer that never shows up must surface as an error after the timeout, not hang the kernel forever
    (signal-pad waits need timeouts)."""
    import time
    from b200ddp.parallel.peer import PeerCollectives, PeerCommError
    dev = torch.device("cuda", rank)
    os.environ["B200DDP_TIMEOUT_S"] = "2"
    comm = PeerCollectives.get(None, dev, min_bytes=8 << 20)
    assert comm.timeout_s == 2.0
    t = torch.ones(1024, device=dev)
    comm.allreduce_([t], wire="fp32", algo="one_shot")            # healthy collective first
    torch.cuda.synchronize()
    comm.check()
    dist.barrier(device_ids=[rank])
    if rank == 0:
        t0 = time.time()
        comm.allreduce_([t], wire="fp32", algo="one_shot")        # rank 1..n never join this one
        torch.cuda.synchronize()
        assert time.time() - t0 < 15
        with pytest.raises(PeerCommError):
            comm.check()
    dist.barrier(device_ids=[rank])


def test_missing_peer_times_out_instead_of_hanging(free_port):
    _spawn(_check_timeout, free_port, world=2)
