"""Autograd bindings for the hand-written sm_100a kernels.

Every op has two bodies: CUDA tensors go to the native extension (and raise if it is missing - no
silent fallback on a GPU box); CPU tensors use plain PyTorch math so the CPU test-suite and the gloo
plumbing config run without a GPU.  This is a device split inside one framework, not a multi-backend
kernel dispatch.

Reference call sites being replaced: ``nn.Linear``/``nn.ReLU`` in ``model.py:11-16`` (cuBLASLt +
clamp kernels), ``nn.MSELoss`` in ``ddp.py:164,222``; LayerNorm / cross-entropy / GELU are needed by the
BERT config named in BASELINE.json.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from .. import _ext

EPI_NONE, EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_GELU = 0, 1, 2, 3
_ACT_TO_EPI = {None: EPI_BIAS, "relu": EPI_BIAS_RELU, "gelu": EPI_BIAS_GELU}


def _C():
    return _ext.get()


# ------------------------------------------------------------------------------------------------
# Linear
# ------------------------------------------------------------------------------------------------
def _tc_ok(M: int, N: int, K: int) -> bool:
    """tcgen05 path needs 16-byte row pitches for TMA for all three GEMMs (fwd, dgrad, wgrad)."""
    return K % 8 == 0 and N % 8 == 0 and M % 8 == 0


class _LinearTC(torch.autograd.Function):
    """bf16 linear on tcgen05: fwd y = act(x W^T + b); bwd dgrad/wgrad without transposes (MN-major
    operand descriptors), ReLU/GELU backward applied to dy before the two GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, activation):
        C = _C()
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        epi = _ACT_TO_EPI[activation] if bias is not None else EPI_NONE
        if bias is None and activation is not None:
            raise ValueError("fused activation needs a bias (use bias=True)")
        if activation == "gelu":
            # keep the pre-activation for the backward (GELU' needs it): two launches, one extra tensor
            pre = C.gemm_nt(x2, weight, bias, EPI_BIAS, None)
            y = C.gelu_fwd(pre)
            ctx.save_for_backward(x2, weight, pre)
        else:
            y = C.gemm_nt(x2, weight, bias, epi, None)
            ctx.save_for_backward(x2, weight, y if activation == "relu" else None)
        ctx.activation = activation
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        C = _C()
        x2, weight, aux = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if ctx.activation == "relu":
            dy2 = dy2 * (aux > 0).to(dy2.dtype)
        elif ctx.activation == "gelu":
            dy2 = C.gelu_bwd(dy2.contiguous(), aux)          # dy * gelu'(pre) in one pass
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[M,K] = dy[M,N] @ W[N,K]   : A = dy (K-major over N), B = W stored [N,K] = "[K_red, N_out]" MN-major
            dx = C.gemm(dy2, weight, None, False, True, EPI_NONE, False, None).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            # dW[N,K] = dy^T[N,M] @ x[M,K] : A = dy stored [M,N] (MN-major), B = x stored [M,K] (MN-major)
            dw = C.gemm(dy2, x2, None, True, True, EPI_NONE, False, None)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db, None


class _LinearSmall(torch.autograd.Function):
    """fp32 linear for shapes below one tensor-core tile (FooModel): CUDA-core kernels, fused
    bias+ReLU forward, single-launch dx/dw/db backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        C = _C()
        xc = x.contiguous()
        y = C.small_linear_fwd(xc, weight.contiguous(), bias, bool(relu))
        ctx.save_for_backward(xc, weight, y)
        ctx.relu = bool(relu)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        C = _C()
        x, weight, y = ctx.saved_tensors
        dx, dw, db = C.small_linear_bwd(dy, x, weight.contiguous(), y, ctx.relu, ctx.needs_input_grad[0], ctx.has_bias)
        return (dx if ctx.needs_input_grad[0] else None), dw, (db if ctx.has_bias else None), None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           activation: Optional[str] = None) -> torch.Tensor:
    """y = act(x @ weight.T + bias), activation in {None, "relu", "gelu"}."""
    if x.is_cuda:
        N, K = weight.shape
        M = x.numel() // K
        if x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and _tc_ok(M, N, K) \
                and (bias is not None or activation is None) and os.environ.get("B200DDP_DISABLE_TC", "0") != "1":
            return _LinearTC.apply(x, weight, bias, activation)
        if x.dtype == torch.float32 and weight.dtype == torch.float32 and N * (K + 1) * 4 <= 48 * 1024 and M * N * 4 <= 48 * 1024 \
                and activation in (None, "relu"):
            return _LinearSmall.apply(x, weight, bias, activation == "relu")
        _C()  # loud failure if the extension is missing; otherwise this shape has no native kernel yet
        y = F.linear(x, weight, bias)
    else:
        y = F.linear(x, weight, bias)
    if activation == "relu":
        y = F.relu(y)
    elif activation == "gelu":
        y = F.gelu(y)
    return y


def conv1x1(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Pointwise (1x1, stride 1, no bias) convolution of a channels_last activation as ONE tcgen05 GEMM:
    the NHWC tensor *is* the row-major [N*H*W, C_in] operand, the [C_out, C_in, 1, 1] filter the [C_out, C_in]
    one, and the [N*H*W, C_out] result is the channels_last output - no im2col, no copies.  dgrad / wgrad reuse
    ``_LinearTC``'s MN-major GEMMs.  Two thirds of ResNet-50's convolutions have this shape."""
    n, c, h, w = x.shape
    co = weight.shape[0]
    ok = (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.shape[2:] == (1, 1)
          and x.is_contiguous(memory_format=torch.channels_last) and _tc_ok(n * h * w, co, c))
    if not ok:
        return F.conv2d(x, weight)
    x2 = x.permute(0, 2, 3, 1).reshape(n * h * w, c)            # view of the NHWC storage
    y2 = _LinearTC.apply(x2, weight.reshape(co, c), None, None)
    return y2.view(n, h, w, co).permute(0, 3, 1, 2)             # channels_last-strided [N, C_out, H, W]


class _Conv1x1Stats(torch.autograd.Function):
    """``_LinearTC`` without bias whose forward GEMM also emits the partial column statistics of its output
    ([2, ceil(M/32), N] fp32: per-32-row sums and sums of squares) for the BatchNorm that follows."""

    @staticmethod
    def forward(ctx, x2, w2):
        y2, partials = _C().gemm_stats(x2, w2)
        ctx.save_for_backward(x2, w2)
        ctx.mark_non_differentiable(partials)
        return y2, partials

    @staticmethod
    def backward(ctx, dy, _dpartials):
        C = _C()
        x2, w2 = ctx.saved_tensors
        dy2 = dy if dy.is_contiguous() else dy.contiguous()
        dx = C.gemm(dy2, w2, None, False, True, EPI_NONE, False, None) if ctx.needs_input_grad[0] else None
        dw = C.gemm(dy2, x2, None, True, True, EPI_NONE, False, None) if ctx.needs_input_grad[1] else None
        return dx, dw


def conv1x1_stats(x: torch.Tensor, weight: torch.Tensor):
    """``conv1x1`` that also returns the BatchNorm partial statistics computed in the GEMM epilogue (or ``None`` when
    the tensor-core path does not apply and the caller should let BatchNorm compute its own)."""
    n, c, h, w = x.shape
    co = weight.shape[0]
    ok = (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.shape[2:] == (1, 1)
          and x.is_contiguous(memory_format=torch.channels_last) and _tc_ok(n * h * w, co, c))
    if not ok:
        return F.conv2d(x, weight), None
    x2 = x.permute(0, 2, 3, 1).reshape(n * h * w, c)
    y2, partials = _Conv1x1Stats.apply(x2, weight.reshape(co, c))
    return y2.view(n, h, w, co).permute(0, 3, 1, 2), partials


# ------------------------------------------------------------------------------------------------
# 3x3 convolution (experimental, opt-in): forward and dgrad on the nine-shifted-GEMM kernel
# ------------------------------------------------------------------------------------------------
def _conv3x3_fprop(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """stride-1 / pad-1 3x3 convolution: ``csrc/conv3x3_tcgen05.cu`` for channels_last bf16 CUDA tensors it covers,
    the stock op otherwise (CPU tests exercise the surrounding autograd logic through this fallback)."""
    if (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.shape[1] % 64 == 0 and w.shape[0] % 8 == 0
            and x.shape[3] <= 128):
        xc = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)
        wc = w if w.is_contiguous(memory_format=torch.channels_last) else w.contiguous(memory_format=torch.channels_last)
        return _C().conv3x3_fwd(xc, wc)
    return F.conv2d(x, w, padding=1)


class _Conv3x3TC(torch.autograd.Function):
    """fprop and dgrad are the same kernel: dx = conv3x3(dy, W') with W'[ci, co, r, s] = W[co, ci, 2-r, 2-s] (a
    few-hundred-KB permutation per layer); wgrad stays on the library (``aten::convolution_backward``) until the split-K
    tcgen05 version exists."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return _conv3x3_fprop(x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            w_t = w.flip(2, 3).transpose(0, 1)
            dx = _conv3x3_fprop(dy, w_t.contiguous(memory_format=torch.channels_last) if w.is_cuda else w_t)
        if ctx.needs_input_grad[1]:
            dw = torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                     [False, True, False])[1]
        return dx, dw


def conv3x3(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """3x3, stride 1, zero padding 1, no bias."""
    return _Conv3x3TC.apply(x, weight)


# ------------------------------------------------------------------------------------------------
# Losses: forward computes loss AND input gradient in one launch
# ------------------------------------------------------------------------------------------------
class _MSEFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        loss, dout = _C().mse_fwd_bwd(out.contiguous(), target.contiguous(), 1.0)
        ctx.save_for_backward(dout)
        return loss.to(out.dtype) if out.dtype != torch.float32 else loss

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g.to(dout.dtype), None


def mse_loss(out: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean((out-target)^2) - reference criterion ``nn.MSELoss`` (``ddp.py:164``)."""
    if out.is_cuda and out.dtype in (torch.float32, torch.bfloat16) and out.dtype == target.dtype and out.shape == target.shape:
        return _MSEFused.apply(out, target)
    return F.mse_loss(out.float(), target.float())


class _XentFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, ignore_index):
        loss, dlogits = _C().xent_fwd_bwd(logits.contiguous(), targets.contiguous(), int(ignore_index), 1.0)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g.to(dlogits.dtype), None, None


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Softmax cross-entropy, mean over non-ignored rows; logits [..., C], targets [...]."""
    if logits.is_cuda and logits.dtype in (torch.float32, torch.bfloat16):
        flat = logits.reshape(-1, logits.shape[-1])
        return _XentFused.apply(flat, targets.reshape(-1), ignore_index)
    return F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), targets.reshape(-1), ignore_index=ignore_index)


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
class _LayerNormFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = x.contiguous()
        y, mean, rstd = _C().layernorm_fwd(xc, gamma.contiguous(), beta.contiguous(), float(eps))
        ctx.save_for_backward(xc, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx, dgamma, dbeta = _C().layernorm_bwd(dy.contiguous(), x, gamma.contiguous(), mean, rstd)
        return dx, dgamma, dbeta, None


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and gamma.dtype == x.dtype and x.shape[-1] * 8 <= 96 * 1024:
        return _LayerNormFused.apply(x, gamma, beta, eps)
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
