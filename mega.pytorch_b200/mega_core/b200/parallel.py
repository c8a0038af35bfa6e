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


# ----------------------------------------------------------------------------------------------------------------
# Wavefront schedule (option ii). The long-range memory couples the key frames of a group, but
# only as a depth-3 wavefront: what frame t pushes into stage s's memory depends on stage s-1's memory of frames < t
# (roi_box_feature_extractors.py:678-688, :913-928), never on its own stage. So the owner of frame t can run the
# whole aggregation of its frame alone if, before stage s reads its memory, it has applied the stage-s increments of
# the group's EARLIER frames (received by one small all-gather per stage) and only those: the ring slots that the
# group's LATER frames will overwrite must still hold the frames they are about to evict, exactly as in the
# sequential order "read, then push". After its last read a rank applies the remaining increments (its own included),
# which leaves every rank with the same ring as sequential processing of the whole group.
def wave_tables(mem_pushed, rank, world, rows0, rows12, mem_frames, base0, base12, baseb12):
    """Destination-row tables of one wavefront step, for the rank that owns frame `rank` of a group of `world` frames.

    mem_pushed : frames pushed into the memory before the group (identical on all ranks)
    rows0 / rows12 : rows per frame of the stage-0 memory (75) / of the stage-1 and stage-2 memories (15)
    mem_frames : ring capacity in frames (25)
    base0 / base12 / baseb12 : first ring row inside the stage-0 feature+box buffers, the stage-1/2 feature buffers and
        the stage-1/2 box buffers
    Returns int32 numpy arrays; entry [g * rows + j] is the destination row of row j of frame g's increment, or -1
    (= skipped by mega_copy_rows): `pre*` hold the frames g < rank, `post*` the frames g >= rank.
    `valid[g]` = memory frames frame g sees (the soft-max's key count is local rows + valid * rows)."""
    import numpy as np
    assert 0 <= rank < world <= mem_frames, "a group must fit the memory ring (its frames take distinct slots)"
    out = {}
    for name, rows, base in (("0", rows0, base0), ("12", rows12, base12), ("b12", rows12, baseb12)):
        pre = np.full(world * rows, -1, dtype=np.int32)
        post = np.full(world * rows, -1, dtype=np.int32)
        for g in range(world):
            slot = (mem_pushed + g) % mem_frames
            dst = base + slot * rows + np.arange(rows, dtype=np.int32)
            (pre if g < rank else post)[g * rows:(g + 1) * rows] = dst
        out["pre" + name], out["post" + name] = pre, post
    out["valid"] = np.asarray([min(mem_pushed + g, mem_frames) for g in range(world)], dtype=np.int32)
    return out


def drive(gen, group=None):
    """run one rank's wavefront generator: every value it yields is (payload, gathered_out); the all-gather result is
    sent back in. Returns the generator's return value."""
    try:
        msg = next(gen)
        while True:
            payload, out = msg
            gather_payloads(payload.view(-1), out, group)
            msg = gen.send(out)
    except StopIteration as stop:
        return stop.value


def play(gens):
    """single-process stand-in for `drive` over all ranks of a group (tests: the ranks of an N-GPU group played on one
    device, or on the CPU for the host logic): advances the generators in lockstep and performs each all-gather by
    copying. Returns the list of return values in rank order."""
    world = len(gens)
    msgs, results, done = [None] * world, [None] * world, [False] * world
    for r, g in enumerate(gens):
        try:
            msgs[r] = next(g)
        except StopIteration as stop:
            results[r], done[r] = stop.value, True
    while not all(done):
        assert not any(done), "ranks left the wavefront at different collectives"
        parts = [m[0].reshape(-1).clone() for m in msgs]
        for r, g in enumerate(gens):
            out = msgs[r][1]
            for q in range(world):
                out[q].copy_(parts[q])
            try:
                msgs[r] = g.send(out)
            except StopIteration as stop:
                results[r], done[r] = stop.value, True
    return results
