"""TrainStep on a GPU: the CUDA-graph step equals the eager step, with and without gradient accumulation
(reference loop: ``ddp.py:227-243`` - loss / accum, backward every micro-step, clip + step + zero_grad on the boundary)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _run(accum, graph, steps=9):
    from b200ddp.engine.step import TrainStep
    from b200ddp.ops import Linear, MSELoss
    from b200ddp.optim import FusedSGD
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = nn.Sequential(Linear(64, 128, activation="relu"), Linear(128, 32)).to(dev)
    opt = FusedSGD(model.parameters(), lr=0.05, max_grad_norm=1.0)
    step = TrainStep(model, MSELoss(), opt, dev, accumulation=accum, use_graph=graph)
    g = torch.Generator().manual_seed(5)
    for i in range(steps * accum):
        x = torch.randn(16, 64, generator=g).to(dev)
        y = torch.randn(16, 32, generator=g).to(dev)
        step(x, y, boundary=(i + 1) % accum == 0)
    torch.cuda.synchronize()
    return [p.detach().clone() for p in model.parameters()], step


@pytest.mark.parametrize("accum", [1, 2, 4])
def test_graph_step_equals_eager_step(accum):
    eager, _ = _run(accum, False)
    graphed, step = _run(accum, True)
    assert step.graph is not None
    assert set(step._graphs) == ({"single"} if accum == 1 else ({"first", "last"} | ({"middle"} if accum > 2 else set())))
    for a, b in zip(eager, graphed):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), float((a - b).abs().max())
    assert abs(step.read_loss_sum()) > 0


def test_boundary_flag_must_follow_the_window():
    from b200ddp.engine.step import TrainStep
    from b200ddp.ops import Linear, MSELoss
    from b200ddp.optim import FusedSGD
    dev = torch.device("cuda", 0)
    model = Linear(8, 8).to(dev)
    step = TrainStep(model, MSELoss(), FusedSGD(model.parameters(), lr=0.1), dev, accumulation=2, use_graph=True, graph_warmup=1)
    x, y = torch.randn(4, 8, device=dev), torch.randn(4, 8, device=dev)
    for i in range(4):
        step(x, y, boundary=i % 2 == 1)
    assert step.graph is not None
    with pytest.raises(RuntimeError, match="boundary"):
        step(x, y, boundary=True)                            # first micro-step of a window cannot be the boundary
