"""Single-process multi-GPU mode (reference ``ddp.py:96-98,189-191``: ``DataParallel(model)`` when launched without a launcher
on a multi-GPU host): forward output and parameter gradients equal the single-device run on the whole batch."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def test_data_parallel_matches_single_device():
    from b200ddp.models import FooModel
    from b200ddp.parallel import DataParallel
    torch.manual_seed(0)
    ref = FooModel().cuda(0)
    dp_inner = FooModel().cuda(0)
    dp_inner.load_state_dict(ref.state_dict())
    dp = DataParallel(dp_inner)
    n = 8 * max(2, torch.cuda.device_count())
    x = torch.randn(n, 10, device="cuda:0")
    y = torch.randn(n, 5, device="cuda:0")
    out_ref = ref(x)
    out_dp = dp(x)
    assert out_dp.device == x.device and torch.allclose(out_dp, out_ref, atol=1e-5)
    torch.nn.functional.mse_loss(out_ref, y).backward()
    torch.nn.functional.mse_loss(out_dp, y).backward()
    for a, b in zip(ref.parameters(), dp_inner.parameters()):
        assert torch.allclose(a.grad, b.grad, atol=1e-5)
