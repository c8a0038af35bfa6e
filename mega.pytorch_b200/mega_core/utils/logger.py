"""`setup_logger(name, save_dir, distributed_rank)` as tools/test_net.py calls it (utils/logger.py of the reference):
rank 0 logs to the console and, when a directory is given, to <save_dir>/log.txt; the other ranks stay silent."""
import logging
import os
import sys

_FORMAT = "%(asctime)s %(name)s %(levelname)s: %(message)s"


def _sinks(save_dir, filename):
    yield logging.StreamHandler(stream=sys.stdout)
    if save_dir:
        yield logging.FileHandler(os.path.join(save_dir, filename))


def setup_logger(name, save_dir, distributed_rank, filename="log.txt"):
    log = logging.getLogger(name)
    log.setLevel(logging.DEBUG)
    if distributed_rank == 0:
        formatter = logging.Formatter(_FORMAT)
        for sink in _sinks(save_dir, filename):
            sink.setLevel(logging.DEBUG)
            sink.setFormatter(formatter)
            log.addHandler(sink)
    return log
