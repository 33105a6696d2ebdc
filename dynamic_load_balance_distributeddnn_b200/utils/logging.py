"""Per-rank logger with the reference's line format and file naming.

Reference ``dbs_logging.py:5-34``: logger → stream + ``./logs/<experiment id>.log`` opened ``w+``,
every record prefixed ``[world_size:lr:dbs_<enabled|disabled>:ft_<enabled|disabled>]``.  Fixed here:
directory creation is race-free (``exist_ok``; the reference's ``os.mkdir`` TOCTOU kills a rank,
SURVEY D4) and the logger is keyed by rank, not hostname, so several ranks in one process (tests)
do not steal each other's handlers.
"""
from __future__ import annotations

import logging
import os

FORMAT = "%(asctime)s [%(world_size)s:%(lr)s:dbs_%(dbs)s:ft_%(ft)s] [%(filename)s:%(lineno)d] %(levelname)s %(message)s"


def init_logger(cfg, rank: int, output_dir: str = None, stream: bool = True) -> logging.LoggerAdapter:
    output_dir = output_dir or cfg.log_dir
    os.makedirs(output_dir, exist_ok=True)
    extra = {"world_size": cfg.world_size, "lr": cfg.learning_rate,
             "dbs": "enabled" if cfg.dynamic_batch_size else "disabled",
             "ft": "enabled" if cfg.fault_tolerance else "disabled"}
    logger = logging.getLogger(f"dlb.rank{rank}.{os.getpid()}")
    for h in logger.handlers[:]:
        logger.removeHandler(h)
        try:
            h.close()
        except Exception:
            pass
    logger.setLevel(logging.DEBUG)
    logger.propagate = False
    fmt = logging.Formatter(FORMAT)
    if stream:
        sh = logging.StreamHandler()
        sh.setLevel(logging.DEBUG if rank == 0 else logging.WARNING)
        sh.setFormatter(fmt)
        logger.addHandler(sh)
    fh = logging.FileHandler(os.path.join(output_dir, cfg.experiment_id(rank) + ".log"), "w+")
    fh.setLevel(logging.DEBUG)
    fh.setFormatter(fmt)
    logger.addHandler(fh)
    return logging.LoggerAdapter(logger, extra)


def log_path(cfg, rank: int) -> str:
    return os.path.join(cfg.log_dir, cfg.experiment_id(rank) + ".log")


def done_marker(cfg) -> str:
    """Completion marker written by rank 0 at the very end.  The reference treats the mere existence of
    rank 0's *log* as "finished" (dbs.py:528-534), so a crashed run is skipped forever."""
    return os.path.join(cfg.log_dir, cfg.experiment_id(0) + ".done")
