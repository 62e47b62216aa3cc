#!/usr/bin/env python
"""Entry point with the reference's launch shape: ``torchrun --nproc_per_node=N ddp.py [flags]`` (or the legacy
``python -m torch.distributed.launch``, or plain ``python ddp.py`` for CPU / single-GPU / DataParallel).
The implementation lives in ``b200ddp.engine``; this file only re-exports the reference's public names:
``set_seed`` (reference ``ddp.py:44-49``), ``get_linear_schedule_with_warmup`` (``ddp.py:52-61``), ``save_model``
(``ddp.py:64-77``), ``setup`` / ``cleanup`` (``ddp.py:80-121``), ``evaluate`` (``ddp.py:123-124``), ``train``
(``ddp.py:126-288``), ``main`` (``ddp.py:291-314``)."""
from b200ddp.engine.cli import build_parser, cleanup, evaluate, main, setup  # noqa: F401
from b200ddp.engine.trainer import Trainer  # noqa: F401
from b200ddp.optim import get_linear_schedule_with_warmup  # noqa: F401
from b200ddp.utils import set_seed  # noqa: F401
from b200ddp.utils.checkpoint import save_model  # noqa: F401


def train(args, model):
    """Reference-shaped helper: ``train(args, model)`` after ``setup(args)``."""
    from b200ddp.engine import cli
    return Trainer(args, model, cli.log).train()


if __name__ == "__main__":
    raise SystemExit(main())
