"""utils/logger.py:7-25: console (+ file) logger for rank 0, silent elsewhere"""
import logging
import os
import sys


def setup_logger(name, save_dir, distributed_rank, filename="log.txt"):
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG)
    if distributed_rank > 0:
        return logger
    fmt = logging.Formatter("%(asctime)s %(name)s %(levelname)s: %(message)s")
    handlers = [logging.StreamHandler(stream=sys.stdout)]
    if save_dir:
        handlers.append(logging.FileHandler(os.path.join(save_dir, filename)))
    for h in handlers:
        h.setLevel(logging.DEBUG)
        h.setFormatter(fmt)
        logger.addHandler(h)
    return logger
