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
