import time

import torch

from b200ddp.utils import StepTimer, nvtx_range, is_dense


def test_step_timer_cpu():
    t = StepTimer(torch.device("cpu"), samples_per_step=64, skip_first=1)
    assert t.summary() is None
    for _ in range(6):
        time.sleep(0.002)
        t.tick()
    s = t.summary()
    assert s["steps"] == 4 and s["ms_per_step"] >= 1.5 and s["samples_per_s"] > 0
    with nvtx_range("noop"):
        pass


def test_is_dense():
    x = torch.randn(2, 3, 4, 5)
    assert is_dense(x) and is_dense(x.contiguous(memory_format=torch.channels_last)) and is_dense(x.permute(3, 0, 1, 2))
    assert not is_dense(x[:, ::2]) and not is_dense(x.expand(2, 2, 3, 4, 5)[0:1].expand(3, 2, 3, 4, 5))
