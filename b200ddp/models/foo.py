"""The reference toy workload (``model.py:8-16``): Linear(10,10) -> ReLU -> Linear(10,5), 165 params.
Here the two linears are ``b200ddp.ops.Linear`` modules - ReLU is fused into the first linear's
epilogue and the backward of each layer is one launch - and parameter names/shapes/init law match the
reference so ``model.bin`` checkpoints are interchangeable (keys ``net1.weight`` ... ``net2.bias``)."""
import torch
import torch.nn as nn

from ..ops import Linear


class FooModel(nn.Module):
    def __init__(self, in_features: int = 10, hidden: int = 10, out_features: int = 5) -> None:
        super().__init__()
        self.net1 = Linear(in_features, hidden, activation="relu")
        self.net2 = Linear(hidden, out_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net2(self.net1(x))


class BranchyFooModel(FooModel):
    """FooModel plus a head that only some inputs exercise - used by the tests for the
    ``find_unused_parameters=True`` semantics the reference always enables (``ddp.py:195``)."""

    def __init__(self) -> None:
        super().__init__()
        self.aux = Linear(10, 5)

    def forward(self, x: torch.Tensor, use_aux: bool = False) -> torch.Tensor:
        h = self.net1(x)
        out = self.net2(h)
        if use_aux:
            out = out + self.aux(h)
        return out
