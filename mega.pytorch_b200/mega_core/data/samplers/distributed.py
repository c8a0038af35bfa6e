"""VIDTestDistributedSampler (data/samplers/distributed.py:69-115): whole videos per rank -- the dataset is cut into
`num_replicas` contiguous index ranges whose borders are moved forward to the next video start, so no video is split
(the reference's "replicas only" multi-GPU mode)."""
import math

import torch
import torch.distributed as dist
from torch.utils.data.sampler import Sampler


class VIDTestDistributedSampler(Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=False):
        self.dataset = dataset
        self.num_replicas = dist.get_world_size() if num_replicas is None else num_replicas
        self.rank = dist.get_rank() if rank is None else rank
        self.epoch, self.shuffle = 0, shuffle
        self.num_samples = int(math.ceil(len(dataset) * 1.0 / self.num_replicas))
        self.start = self._next_video_start(self.rank * self.num_samples)
        self.end = self._next_video_start((self.rank + 1) * self.num_samples)

    def _next_video_start(self, offset):
        if offset >= len(self.dataset):
            return len(self.dataset)
        for index in self.dataset.start_index:
            if index >= offset:
                return index
        return None          # like the reference: no later video start -> slice to the end

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.epoch)
            indices = torch.randperm(len(self.dataset), generator=g).tolist()
        else:
            indices = list(range(len(self.dataset)))
        return iter(indices[self.start:self.end])

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch
