"""Reference-named helpers re-exported from the package: ``getLoggerWithRank`` (reference ``utils.py:9-75``),
warnings redirection (``utils.py:78-82``), rank helpers (``utils.py:84-101``)."""
from b200ddp.utils import (getLoggerWithRank, get_logger_with_rank, redirect_warnings_to_logger,  # noqa: F401
                           get_rank, get_world_size, is_main_process)
