from __future__ import annotations

import logging
import sys

LOGGER_NAME = "d9d"


def build_dist_logger(qualifier: str, level: int) -> logging.Logger:
    """Process-wide ``d9d`` logger writing to stdout, prefixed with this rank's mesh coordinates."""
    logger = logging.getLogger(LOGGER_NAME)
    logger.setLevel(level)
    for handler in list(logger.handlers):
        logger.removeHandler(handler)
    handler = logging.StreamHandler(sys.stdout)
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter(f"[d9d] [{qualifier}] %(asctime)s - %(levelname)s - %(message)s"))
    logger.addHandler(handler)
    logger.propagate = False
    return logger
