"""Grouped tile rasterisation (B200DDP_GEMM_GROUP_M / set_gemm_group_m): results must be bit-identical to the
default m-fastest order, since only the order in which persistent CTAs pick tiles changes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K", [(2048, 2048, 512), (1000, 3000, 264), (4096, 768, 768)])
def test_grouped_raster_is_bit_identical(mode, M, N, K):
    from b200ddp import _ext
    C = _ext.get()
    torch.manual_seed(1)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    try:
        C.set_gemm_cta_mode(mode)
        C.set_gemm_group_m(0)
        d0 = C.gemm(a, b, None, False, False, 0, False, None)
        outs = []
        for g in (1, 4, 8):
            C.set_gemm_group_m(g)
            outs.append(C.gemm(a, b, None, False, False, 0, False, None))
    finally:
        C.set_gemm_group_m(0)
        C.set_gemm_cta_mode(0)
    ref = a.float() @ b.float().t()
    assert float((d0.float() - ref).norm() / ref.norm()) < 1e-2
    for d in outs:
        assert torch.equal(d, d0)
