"""Opt-in (default-off) paths prepared for measurement: zero-padded stem input, 1x1 convolutions on the tcgen05 GEMM.
Sorted last on purpose: these paths are not in the default step."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("B200DDP_TEST_OPTIN") != "1",
                                                   reason="written after the last GPU session: run with B200DDP_TEST_OPTIN=1 (tools/round2_ablation.sh)")]


def test_normalize_pads_channels_with_zeros():
    from b200ddp import _ext
    C = _ext.get()
    x = torch.randn(4, 3, 32, 40, device="cuda")
    mean = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    istd = torch.tensor([2.0, 0.5, 1.5], device="cuda")
    dst = torch.full((4, 8, 32, 40), 7.0, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    C.normalize_to_channels_last(x, dst, mean, istd, 1.0)
    ref = ((x - mean.view(1, 3, 1, 1)) * istd.view(1, 3, 1, 1)).to(torch.bfloat16)
    assert torch.allclose(dst[:, :3].float(), ref.float(), atol=1e-2, rtol=1e-2)
    assert float(dst[:, 3:].abs().max()) == 0.0


def test_resnet_padded_stem_matches_default_on_gpu():
    from b200ddp.models.resnet import ResNet
    from b200ddp.utils import to_mixed_bf16
    torch.manual_seed(0)
    a = to_mixed_bf16(ResNet([1, 1, 1, 1], num_classes=16).cuda()).to(memory_format=torch.channels_last)
    b = to_mixed_bf16(ResNet([1, 1, 1, 1], num_classes=16, stem_pad_to=8).cuda()).to(memory_format=torch.channels_last)
    b.load_state_dict(a.state_dict())
    x = torch.randn(8, 3, 64, 64, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ya, yb = a(x), b(x)
    assert float((ya.float() - yb.float()).norm() / ya.float().norm()) < 3e-2
    ya.float().square().mean().backward()
    yb.float().square().mean().backward()
    ga, gb = a.conv1.weight.grad.float(), b.conv1.weight.grad.float()
    assert gb.shape == ga.shape
    assert float((ga - gb).norm() / ga.norm()) < 5e-2


@pytest.mark.parametrize("shape", [(8, 64, 56, 56, 256), (4, 256, 14, 14, 64), (2, 512, 7, 7, 2048)])
def test_conv1x1_on_tcgen05_matches_cudnn(shape):
    from b200ddp.ops import PointwiseConv2d
    n, ci, h, w, co = shape
    torch.manual_seed(0)
    ref = PointwiseConv2d(ci, co, use_tc=False).cuda().bfloat16().to(memory_format=torch.channels_last)
    tc = PointwiseConv2d(ci, co, use_tc=True).cuda().bfloat16().to(memory_format=torch.channels_last)
    tc.load_state_dict(ref.state_dict())
    x = torch.randn(n, ci, h, w, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), tc(xb)
    assert yb.shape == ya.shape and yb.is_contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))
    assert rel(yb, ya) < 1e-2
    assert rel(xb.grad, xa.grad) < 1e-2
    assert tc.weight.grad.shape == ref.weight.grad.shape
    assert rel(tc.weight.grad, ref.weight.grad) < 1e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="DataParallel needs two GPUs in one process")
def test_data_parallel_matches_single_device():
    """Reference ``ddp.py:189-191`` mode: one process, several GPUs.  Same loss and gradients as one device."""
    from b200ddp.models import FooModel
    from b200ddp.parallel import DataParallel
    torch.manual_seed(0)
    single = FooModel().cuda(0)
    multi = FooModel().cuda(0)
    multi.load_state_dict(single.state_dict())
    dp = DataParallel(multi, device_ids=[0, 1])
    x = torch.randn(64, 10, device="cuda:0")
    y = torch.randn(64, 5, device="cuda:0")
    la = torch.nn.functional.mse_loss(single(x), y)
    lb = torch.nn.functional.mse_loss(dp(x), y)
    la.backward()
    lb.backward()
    assert torch.allclose(la, lb, atol=1e-6)
    for p, q in zip(single.parameters(), multi.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5)
    assert set(dp.state_dict()) == set(single.state_dict())        # no "module." prefix in checkpoints


@pytest.mark.parametrize("mode,tma", [(1, 0), (2, 0), (1, 1), (2, 1)])
@pytest.mark.parametrize("M,N,K", [(1000, 264, 72), (25088, 256, 64), (6272, 1024, 256), (304, 64, 512)])
def test_gemm_epilogue_column_statistics(mode, tma, M, N, K):
    """Per-32-row partial column sums / sums of squares written by the GEMM epilogue == the same sums of its stored output."""
    from b200ddp import _ext
    C = _ext.get()
    torch.manual_seed(3)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.2
    try:
        C.set_gemm_cta_mode(mode)
        C.set_gemm_tma_store(tma)
        d, st = C.gemm_stats(a, b)
        d0 = C.gemm(a, b, None, False, False, 0, False, None)
    finally:
        C.set_gemm_tma_store(0)
        C.set_gemm_cta_mode(0)
    assert torch.equal(d, d0)
    G = (M + 31) // 32
    assert st.shape == (2, G, N)
    pad = torch.zeros(G * 32, N, device="cuda")
    pad[:M] = d.float()
    grp = pad.view(G, 32, N)
    assert torch.allclose(st[0], grp.sum(1), atol=1e-3, rtol=1e-4)
    assert torch.allclose(st[1], grp.square().sum(1), atol=1e-2, rtol=1e-4)


def test_bottleneck_with_stats_from_the_gemm_epilogue(monkeypatch):
    from b200ddp.models.resnet import Bottleneck
    from b200ddp.utils import to_mixed_bf16
    monkeypatch.setenv("B200DDP_CONV1X1_TC", "1")
    torch.manual_seed(0)
    monkeypatch.setenv("B200DDP_CONV_BN_FUSE", "0")
    ref = to_mixed_bf16(Bottleneck(256, 64).cuda()).to(memory_format=torch.channels_last)
    monkeypatch.setenv("B200DDP_CONV_BN_FUSE", "1")
    fused = to_mixed_bf16(Bottleneck(256, 64).cuda()).to(memory_format=torch.channels_last)
    assert fused.fuse_stats and not ref.fuse_stats
    fused.load_state_dict(ref.state_dict())
    x = torch.randn(16, 256, 28, 28, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), fused(xb)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)

    def rel(u, v):
        return float((u.float() - v.float()).norm() / (v.float().norm() + 1e-12))
    assert rel(yb, ya) < 2e-2
    assert rel(xb.grad, xa.grad) < 3e-2
    for (k, p), (_, q) in zip(ref.named_parameters(), fused.named_parameters()):
        if p.grad.float().norm() > 1e-3:
            assert rel(q.grad, p.grad) < 5e-2, k
    for (k, u), (_, v) in zip(ref.named_buffers(), fused.named_buffers()):
        assert torch.allclose(u.float(), v.float(), atol=1e-2, rtol=1e-2), k


@pytest.mark.parametrize("N,C,H,W,K", [(32, 64, 56, 56, 64), (32, 128, 28, 28, 128), (32, 256, 14, 14, 256), (32, 512, 7, 7, 512),
                                       (3, 64, 7, 7, 24), (4, 64, 12, 20, 72)])
def test_conv3x3_nine_shifted_gemms_matches_cudnn(N, C, H, W, K):
    """Experimental tcgen05 3x3 convolution forward (csrc/conv3x3_tcgen05.cu) vs F.conv2d."""
    from b200ddp import _ext
    Cx = _ext.get()
    torch.manual_seed(0)
    x = torch.randn(N, C, H, W, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, 3, 3, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = Cx.conv3x3_fwd(x, w)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), padding=1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    err = float((y.float() - ref).norm() / ref.norm())
    assert err < 1e-2, err


def test_conv3x3_module_forward_and_dgrad_on_the_draft_kernel():
    from b200ddp.ops import Conv3x3
    torch.manual_seed(0)
    ref = Conv3x3(128, 128, use_tc=False).cuda().bfloat16().to(memory_format=torch.channels_last)
    tc = Conv3x3(128, 128, use_tc=True).cuda().bfloat16().to(memory_format=torch.channels_last)
    tc.load_state_dict(ref.state_dict())
    x = torch.randn(8, 128, 28, 28, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), tc(xb)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)

    def rel(u, v):
        return float((u.float() - v.float()).norm() / (v.float().norm() + 1e-12))
    assert rel(yb, ya) < 1e-2 and rel(xb.grad, xa.grad) < 1e-2 and rel(tc.weight.grad, ref.weight.grad) < 1e-2


@pytest.mark.parametrize("shape", [(32, 64, 56, 56), (32, 1024, 14, 14), (8, 2048, 7, 7)])
def test_batchnorm_programmatic_dependent_launch_is_bit_identical(shape):
    """B200DDP_PDL: stats -> apply and bwd-reduce -> bwd-apply as programmatic dependent launches: same kernels' bodies,
    same bits, eagerly and inside a captured CUDA graph."""
    from b200ddp import _ext
    from b200ddp.ops import FusedBatchNormAct2d
    C = _ext.get()
    torch.manual_seed(0)
    x = torch.randn(*shape, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = torch.randn_like(x)
    g = torch.randn_like(x)

    def run(pdl):
        C.set_bn_pdl(pdl)
        bn = FusedBatchNormAct2d(shape[1], relu=True).cuda()
        xa = x.clone().requires_grad_(True)
        y = bn(xa, residual=res)
        y.backward(g)
        return y.detach().clone(), xa.grad.clone(), bn.weight.grad.clone(), bn.running_var.clone()
    try:
        base = run(0)
        pdl = run(1)
        for a, b in zip(base, pdl):
            assert torch.equal(a, b)
        # under stream capture the dependent launch becomes a programmatic edge of the graph
        bn = FusedBatchNormAct2d(shape[1], relu=True).cuda()
        bn(x)                                             # allocate workspaces outside the capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = bn(x)
        graph.replay()
        torch.cuda.synchronize()
        C.set_bn_pdl(0)
        ref_bn = FusedBatchNormAct2d(shape[1], relu=True).cuda()
        assert torch.equal(out, ref_bn(x))
    finally:
        C.set_bn_pdl(0)
