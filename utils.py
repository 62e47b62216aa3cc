"""Reference-named helpers (``utils.py`` in the reference) re-exported from the package."""
from b200ddp.utils import (getLoggerWithRank, get_logger_with_rank, redirect_warnings_to_logger,  # noqa: F401
                           get_rank, get_world_size, is_main_process)
