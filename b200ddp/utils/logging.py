"""Rank-aware structured logging.

Behavioural parity target: reference ``utils.py:9-82`` (line shape, key=value
suffixes from a mapping passed as the sole %-arg, tqdm-safe emission, INFO on
local rank -1/0 and WARNING elsewhere, ``warnings`` rerouted to the logger).
The implementation is independent: one formatter assembles the whole line,
rank fields travel on a ``LoggerAdapter``-free filter object, and the
reference's "TODO: Add File Handler" (``utils.py:61``) is implemented via
``log_file=``.
"""
from __future__ import annotations

import logging
import sys
import time
import warnings
from collections.abc import Mapping
from typing import Optional

try:  # tqdm keeps progress bars intact when log lines are written through it
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover - tqdm is present in the image
    _tqdm = None

MAIN_LOCAL_RANKS = (-1, 0)


class RankLineFormatter(logging.Formatter):
    """``[time] [LEVEL   ] [rank ^ local] [module] [file:line] [msg] [k=v]...``"""

    def __init__(self, node_rank: int = -1, local_rank: int = -1):
        super().__init__()
        self.node_rank = node_rank
        self.local_rank = local_rank

    @staticmethod
    def _stamp(created: float) -> str:
        whole = int(created)
        millis = int((created - whole) * 1000)
        lt = time.localtime(whole)
        return "%s.%03d%s" % (time.strftime("%Y-%m-%d %H:%M:%S", lt), millis, time.strftime("%z", lt))

    def format(self, record: logging.LogRecord) -> str:
        fields = record.args if isinstance(record.args, Mapping) else None
        if fields is not None:
            # the mapping is metadata, not %-format input
            message = str(record.msg)
        else:
            message = record.getMessage()
        node_rank = getattr(record, "node_rank", self.node_rank)
        local_rank = getattr(record, "local_rank", self.local_rank)
        parts = [
            "[%s]" % self._stamp(record.created),
            "[%-8s]" % record.levelname,
            "[%d ^ %d]" % (node_rank, local_rank),
            "[%s]" % record.module,
            "[%s:%d]" % (record.filename, record.lineno),
            "[%s]" % message,
        ]
        if fields is not None:
            parts.extend("[%s=%r]" % (key, fields[key]) for key in fields)
        line = " ".join(parts)
        if record.exc_info:
            line += "\n" + self.formatException(record.exc_info)
        return line


class ProgressSafeHandler(logging.Handler):
    """Writes to stdout through ``tqdm.write`` so bars are not torn."""

    def emit(self, record: logging.LogRecord) -> None:
        try:
            text = self.format(record)
            if _tqdm is not None:
                _tqdm.write(text, file=sys.stdout)
            else:
                print(text, file=sys.stdout)
            sys.stdout.flush()
        except (KeyboardInterrupt, SystemExit):
            raise
        except Exception:
            self.handleError(record)


class _RankStamp(logging.Filter):
    def __init__(self, node_rank: int, local_rank: int):
        super().__init__()
        self.node_rank, self.local_rank = node_rank, local_rank

    def filter(self, record: logging.LogRecord) -> bool:
        record.node_rank = self.node_rank
        record.local_rank = self.local_rank
        return True


def get_logger_with_rank(name: str, node_rank: int, local_rank: int,
                         log_file: Optional[str] = None) -> logging.Logger:
    """Build (or rebuild) the per-process logger.

    Non-main local ranks are capped at WARNING so "every rank must see this"
    events are sent with ``log.warning`` (same convention as the reference,
    ``utils.py:67-68``).
    """
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO if local_rank in MAIN_LOCAL_RANKS else logging.WARNING)
    for old in list(logger.handlers):
        logger.removeHandler(old)
    for old in list(logger.filters):
        logger.removeFilter(old)
    fmt = RankLineFormatter(node_rank, local_rank)
    console = ProgressSafeHandler()
    console.setFormatter(fmt)
    logger.addHandler(console)
    if log_file:
        fh = logging.FileHandler(log_file.format(rank=node_rank, local_rank=local_rank))
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    logger.addFilter(_RankStamp(node_rank, local_rank))
    logger.propagate = False
    return logger


# reference spelling
getLoggerWithRank = get_logger_with_rank


def redirect_warnings_to_logger(logger: logging.Logger) -> None:
    """Route ``warnings.warn`` output into ``logger.warning`` with file/line fields."""

    def _show(message, category, filename, lineno, file=None, line=None):
        logger.warning(message, {"filename": filename, "lineno": lineno})

    warnings.showwarning = _show
