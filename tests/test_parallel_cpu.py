"""world_size-2 gloo test of the frame-parallel host logic (mega_core/b200/parallel.py): payloads are
gathered in frame order and every rank sees the identical sequence."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "mega.pytorch_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "mega_parallel", os.path.join(root, "mega.pytorch_b200", "mega_core", "b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    seen = []
    for step in range(3):
        frames = par.frames_of_step(step, world)
        mine = [f for f in frames if par.owner_of(f, world) == rank]
        assert len(mine) == 1
        payload = torch.full((1028,), float(mine[0]))           # stands for the 75x1028 ROI payload of frame `mine`
        payload[1] = rank
        allp = par.gather_payloads(payload)
        assert allp.shape == (world, 1028)
        for g, f in enumerate(frames):                           # ingestion order == frame order on every rank
            assert allp[g, 0].item() == f and allp[g, 1].item() == par.owner_of(f, world)
            seen.append(int(allp[g, 0].item()))
    q.put((rank, seen))
    dist.destroy_process_group()


def test_frame_parallel_gather_order_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] == list(range(3 * world))


def _pred_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "mega.pytorch_b200"))
    from mega_core.structures.bounding_box import BoxList
    from mega_core.utils import comm
    g = torch.Generator().manual_seed(5)
    everything = {}
    for i in range(11):                                  # identical on both ranks; each keeps its shard (i % world)
        n = int(torch.randint(0, 7, (1,), generator=g))
        b = BoxList(torch.rand(n, 4, generator=g) * 100, (640, 360 + i), mode="xyxy")
        b.add_field("scores", torch.rand(n, generator=g))
        b.add_field("labels", torch.randint(1, 31, (n,), generator=g))
        everything[i] = b
    mine = {i: b for i, b in everything.items() if i % world == rank}
    out = comm.gather_predictions(mine)
    ok = True
    if rank == 0:
        ok = len(out) == 11
        for i, b in enumerate(out):
            e = everything[i]
            ok = ok and b.size == e.size and torch.equal(b.bbox, e.bbox) and torch.equal(b.get_field("scores"), e.get_field("scores")) \
                and torch.equal(b.get_field("labels"), e.get_field("labels"))
    else:
        ok = out is None
    objs = comm.all_gather({"rank": rank, "n": len(mine)})
    ok = ok and [o["rank"] for o in objs] == list(range(world)) and comm.get_world_size() == world
    comm.synchronize()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_packed_prediction_gather_gloo():
    """mega_core.utils.comm.gather_predictions: rank 0 receives every rank's detections as
    typed tensors and rebuilds the list of BoxLists ordered by image id, exactly the per-rank inputs"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pred_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
