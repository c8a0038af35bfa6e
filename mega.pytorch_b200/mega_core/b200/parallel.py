"""Frame-parallel sharding of one video stream over `world` GPUs.

Key frame t of a group of `world` consecutive key frames is owned by rank t mod world: the owner runs the
per-frame branch of the (look-ahead local, global) frame pair that arrives with it; the fixed-size
payloads are exchanged with one all-gather whose output order IS the frame order, so every rank then
ingests the same frames in the same order and all replicas hold identical state (no reductions:
results are independent of `world`)."""
import torch
import torch.distributed as dist


def owner_of(key_frame, world):
    return key_frame % world


def frames_of_step(step, world):
    """key frames produced by distributed step `step` (in ingestion order)"""
    return list(range(step * world, (step + 1) * world))


def gather_payloads(payload, out=None, group=None):
    """all-gather equal-size 1-D payloads -> [world, n] in rank (= frame) order. NCCL on GPUs
    (`all_gather_into_tensor`), gloo for the CPU tests of this host logic."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world, payload.numel(), dtype=payload.dtype, device=payload.device)
    if payload.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), payload.contiguous(), group=group)
    else:
        parts = [out[i] for i in range(world)]
        dist.all_gather(parts, payload.contiguous(), group=group)
    return out
