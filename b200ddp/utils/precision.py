"""Mixed-precision placement: bf16 weights/activations with BatchNorm kept in fp32 (cuDNN's fast NHWC
batch-norm kernels want bf16 activations with fp32 scale/bias/statistics).  This is the working version of
what the reference's ``--fp16`` asks apex for (O2: half model + fp32 master weights, ``ddp.py:174-180``); the
fp32 masters live in ``b200ddp.optim.FusedSGD``."""
from __future__ import annotations

import torch
import torch.nn as nn

_KEEP_FP32 = (nn.modules.batchnorm._BatchNorm,)


def to_mixed_bf16(model: nn.Module, keep_norm_fp32: bool = True) -> nn.Module:
    model.to(torch.bfloat16)
    if keep_norm_fp32:
        for m in model.modules():
            if isinstance(m, _KEEP_FP32):
                m.float()
    return model


def is_dense(t: torch.Tensor) -> bool:
    """True when the tensor's elements occupy one gap-free block of storage (any permutation of strides)."""
    if t.is_contiguous():
        return True
    try:
        from torch._prims_common import is_non_overlapping_and_dense
        return bool(is_non_overlapping_and_dense(t))
    except Exception:
        if t.dim() == 4:
            return t.is_contiguous(memory_format=torch.channels_last)
        if t.dim() == 5:
            return t.is_contiguous(memory_format=torch.channels_last_3d)
        return False
