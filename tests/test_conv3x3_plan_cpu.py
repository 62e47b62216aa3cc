"""Host-side checks for the experimental tcgen05 3x3 convolution (csrc/conv3x3_tcgen05.cu).

The kernel itself needs a GPU (tests/test_gpu_zz_optin.py); what can be pinned down here is its *decomposition*: patch
shapes, tap -> TMA coordinates, filter column order, accumulator row -> pixel.  ``emulate`` below walks exactly the
indices the kernel walks (same names), with zero fill standing in for the TMA unit's out-of-bounds behaviour."""
import pytest
import torch
import torch.nn.functional as F

from b200ddp import _ext

pytestmark = pytest.mark.skipif(not _ext.available(), reason="native extension not built")


@pytest.mark.parametrize("N,H,W,expect", [(32, 56, 56, (2, 1)), (32, 28, 28, (4, 1)), (32, 14, 14, (7, 1)), (32, 7, 7, (7, 2)),
                                          (3, 7, 7, (7, 1)), (4, 5, 5, (5, 4)), (2, 9, 128, (1, 1))])
def test_patch_shapes(N, H, W, expect):
    ok, bh, bi = _ext.get().conv3x3_patch(N, H, W)
    assert ok and (bh, bi) == expect
    assert H % bh == 0 and N % bi == 0 and W * bh * bi <= 128


def test_rows_wider_than_a_tile_are_rejected():
    assert not _ext.get().conv3x3_patch(1, 8, 129)[0]


def emulate(x, w):
    """x [N,C,H,W], w [K,C,3,3] (fp32 here).  Mirrors conv3x3_fprop_kernel's loops."""
    N, C, H, W = x.shape
    K = w.shape[0]
    ok, BH, BI = _ext.get().conv3x3_patch(N, H, W)
    assert ok
    xn = x.permute(0, 2, 3, 1).contiguous()                       # NHWC
    wk = w.permute(0, 2, 3, 1).reshape(K, 9 * C)                  # [K][r][s][C] == row-major [K, 9C]
    y = torch.zeros(N, H, W, K)
    tiles_h, tiles_n = H // BH, N // BI
    patch_rows = W * BH * BI
    for mt in range(tiles_n * tiles_h):
        img0, h0 = (mt // tiles_h) * BI, (mt % tiles_h) * BH
        acc = torch.zeros(128, K)
        for tap in range(9):
            r, s = tap // 3, tap % 3
            for cb in range(C // 64):
                a = torch.zeros(128, 64)                          # smem A tile: rows in box order (w fastest, then h, then image)
                for bi in range(BI):
                    for hh in range(BH):
                        for ww in range(W):
                            hs, ws_ = h0 + hh + r - 1, ww + s - 1  # TMA coordinates (c0, s-1, h0+r-1, img0) + box offsets
                            if 0 <= hs < H and 0 <= ws_ < W:       # outside the tensor: zero fill
                                a[(bi * BH + hh) * W + ww] = xn[img0 + bi, hs, ws_, cb * 64:(cb + 1) * 64]
                b = wk[:, tap * C + cb * 64: tap * C + (cb + 1) * 64]   # B tile: box at column tap*C + cb*64
                acc += a @ b.t()
        for i in range(patch_rows):                               # epilogue: accumulator row -> pixel
            bi, rem = divmod(i, W * BH)
            hh, ww = divmod(rem, W)
            y[img0 + bi, h0 + hh, ww] = acc[i]
    return y.permute(0, 3, 1, 2)


@pytest.mark.parametrize("N,C,H,W,K", [(2, 64, 8, 8, 16), (4, 64, 7, 7, 8), (1, 128, 6, 10, 24)])
def test_nine_shifted_gemms_equal_the_convolution(N, C, H, W, K):
    torch.manual_seed(0)
    x = torch.randn(N, C, H, W)
    w = torch.randn(K, C, 3, 3) * 0.1
    ref = F.conv2d(x, w, padding=1)
    assert torch.allclose(emulate(x, w), ref, atol=1e-3, rtol=1e-4)


def test_conv3x3_function_gradients_match_autograd():
    """dgrad as a forward convolution with the mirrored, transposed filter; wgrad from aten::convolution_backward."""
    from b200ddp.ops import Conv3x3, conv3x3
    torch.manual_seed(0)
    x = torch.randn(2, 8, 9, 7, requires_grad=True)
    w = (torch.randn(6, 8, 3, 3) * 0.2).requires_grad_(True)
    x2, w2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    g = torch.randn(2, 6, 9, 7)
    conv3x3(x, w).backward(g)
    F.conv2d(x2, w2, padding=1).backward(g)
    assert torch.allclose(x.grad, x2.grad, atol=1e-4)
    assert torch.allclose(w.grad, w2.grad, atol=1e-4)
    m = Conv3x3(8, 6, use_tc=True)
    ref = torch.nn.Conv2d(8, 6, 3, padding=1, bias=False)
    ref.load_state_dict(m.state_dict())
    assert torch.allclose(m(x2.detach()), ref(x2.detach()), atol=1e-5)
    assert not Conv3x3(8, 6, stride=2, use_tc=True).use_tc        # strided instances stay on the library
