"""utils/dist_env.py:7-48: process-group start for `--launcher pytorch` (torchrun / torch.distributed.launch:
RANK, WORLD_SIZE, LOCAL_RANK, MASTER_* in the environment), one process per GPU over NCCL."""
import os

import torch
import torch.distributed as dist


def init_dist(launcher, args=None, backend="nccl"):
    if launcher != "pytorch":
        raise ValueError("mega_core (B200 build): launcher %r is not supported, use torchrun" % launcher)
    local_rank = int(os.environ.get("LOCAL_RANK", getattr(args, "local_rank", 0) or 0))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    dist.init_process_group(backend=backend, init_method="env://")
