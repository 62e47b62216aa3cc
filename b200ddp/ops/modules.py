"""``nn.Module`` wrappers over ``ops.functional`` (drop-in for nn.Linear / nn.LayerNorm / loss modules).
Reference call sites: ``nn.Linear`` x2 + ``nn.ReLU`` in ``model.py:11-16``, ``nn.MSELoss`` in ``ddp.py:164``."""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn


class Linear(nn.Module):
    """``nn.Linear`` replacement with an optional fused activation epilogue."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, activation: Optional[str] = None,
                 device=None, dtype=None):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = nn.Parameter(torch.empty(out_features, in_features, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.empty(out_features, device=device, dtype=dtype)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        # same init law as nn.Linear so seeds give the same starting point as the reference model
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.linear(x, self.weight, self.bias, self.activation)

    def extra_repr(self) -> str:
        return f"in={self.in_features}, out={self.out_features}, bias={self.bias is not None}, act={self.activation}"


class PointwiseConv2d(nn.Conv2d):
    """1x1 bias-free ``nn.Conv2d`` (same parameter name, shape and init).  Default: cuDNN, exactly like
    ``nn.Conv2d``.  Opt-in (``B200DDP_CONV1X1_TC=1`` at construction, or ``use_tc=True``): stride-1 instances
    run fprop / dgrad / wgrad on the tcgen05 GEMM (``functional.conv1x1``)."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, use_tc: Optional[bool] = None, **kw):
        super().__init__(in_channels, out_channels, 1, stride=stride, bias=False, **kw)
        if use_tc is None:
            use_tc = os.environ.get("B200DDP_CONV1X1_TC", "0") == "1"
        self.use_tc = bool(use_tc) and self.stride == (1, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.use_tc and x.is_cuda:
            return Fn.conv1x1(x, self.weight)
        return super().forward(x)

    def forward_with_stats(self, x: torch.Tensor):
        """(y, partial BatchNorm statistics from the GEMM epilogue) - statistics are ``None`` off the tensor-core path."""
        if self.use_tc and x.is_cuda:
            return Fn.conv1x1_stats(x, self.weight)
        return super().forward(x), None


class Conv3x3(nn.Conv2d):
    """3x3 / pad 1 bias-free ``nn.Conv2d`` (same parameter name, shape, init).  Default: cuDNN.  Opt-in
    (``B200DDP_CONV3X3_TC=1`` at construction, or ``use_tc=True``): stride-1 instances run forward and dgrad on the
    experimental nine-shifted-GEMM tcgen05 kernel (``functional.conv3x3``)."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, use_tc: Optional[bool] = None, **kw):
        super().__init__(in_channels, out_channels, 3, stride=stride, padding=1, bias=False, **kw)
        if use_tc is None:
            use_tc = os.environ.get("B200DDP_CONV3X3_TC", "0") == "1"
        self.use_tc = bool(use_tc) and self.stride == (1, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.use_tc:
            return Fn.conv3x3(x, self.weight)
        return super().forward(x)


class LayerNorm(nn.Module):
    def __init__(self, hidden: int, eps: float = 1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(hidden, device=device, dtype=dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.layer_norm(x, self.weight, self.bias, self.eps)


class MSELoss(nn.Module):
    def forward(self, out: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return Fn.mse_loss(out, target)


class CrossEntropyLoss(nn.Module):
    def __init__(self, ignore_index: int = -100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return Fn.cross_entropy(logits, target, self.ignore_index)
