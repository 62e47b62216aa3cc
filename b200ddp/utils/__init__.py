from .logging import (get_logger_with_rank, getLoggerWithRank, redirect_warnings_to_logger,
                      RankLineFormatter, ProgressSafeHandler)
from .dist import (get_rank, get_world_size, is_main_process, resolve_local_rank, env_int,
                   barrier_all, max_over_ranks)
from .seed import set_seed, rng_state, restore_rng_state
from .precision import to_mixed_bf16, is_dense
from .timing import StepTimer, nvtx_range

__all__ = [
    "get_logger_with_rank", "getLoggerWithRank", "redirect_warnings_to_logger", "RankLineFormatter",
    "ProgressSafeHandler", "get_rank", "get_world_size", "is_main_process", "resolve_local_rank",
    "env_int", "barrier_all", "max_over_ranks", "set_seed", "rng_state", "restore_rng_state", "to_mixed_bf16", "is_dense", "StepTimer", "nvtx_range",
]
