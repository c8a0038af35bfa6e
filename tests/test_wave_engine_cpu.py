"""MegaEngine's wavefront multi-GPU step executed on the CPU stand-ins (tests/cpu_ops.py): the engine's real host code
-- _wave generator, destination tables, payload ingestion order, _wave_a.._wave_d launch sequences, parallel.play --
runs for groups of 2 and 3 ranks and must reproduce, BIT for bit, the detections, predictor outputs and every memory /
window ring of the sequential owner-mode step (dist_step with world = 1), including across ring wrap-around.
(The CUDA kernels are not involved; their parity tests are the -m gpu suite.)"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpu_ops import cpu_ops  # noqa: E402

W_IMG, H_IMG = 320, 192


def _make(sd, precision="tf32"):
    from mega_core.b200 import engine
    cfg = engine.EngineConfig(precision=precision, all_frame_interval=5, key_frame_location=2, memory_size=4, global_size=2,
                              post_nms_top_n=24, ref_post_nms_top_n=10)          # R = 10, A = 2, KP = 24
    eng = engine.MegaEngine(sd, cfg, device="cpu")
    eng.use_graph = False
    return eng


def _boxes(g, n):
    x1 = torch.rand(n, generator=g) * (W_IMG - 60)
    y1 = torch.rand(n, generator=g) * (H_IMG - 60)
    return torch.stack([x1, y1, x1 + 8 + torch.rand(n, generator=g) * 50, y1 + 8 + torch.rand(n, generator=g) * 50], 1)


def _prime(eng, seed):
    """the state a video is in once its window and global pool are full (what start_video leaves behind)"""
    g = torch.Generator().manual_seed(seed)
    eng.reset()
    for _ in range(eng.L):
        eng._claim_slot()
    eng.win_x.copy_(torch.randn(eng.win_x.shape, generator=g) * 0.5)
    eng.win_boxes.copy_(_boxes(g, eng.win_boxes.shape[0]))
    eng.win_cnt.fill_(eng.KP - 3)
    eng.glob_x.copy_(torch.randn(eng.glob_x.shape, generator=g) * 0.5)
    eng.glob_pushed = eng.GF


def _payload(eng, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros_like(eng.payload_in)
    px, pb, pc, pg = eng._payload_views(p)
    px.copy_(torch.randn(px.shape, generator=g) * 0.5)
    pb.copy_(_boxes(g, pb.shape[0]))
    pc.view(torch.int32)[0, 0] = eng.KP - 1 - seed % 4
    pg.copy_(torch.randn(pg.shape, generator=g) * 0.5)
    return p


def _snap(eng, det, k=None):
    k = int(eng.cur_cnt.view(-1)[0]) if k is None else k
    n = int(det.count.reshape(-1)[0])
    return eng.last_pred[:k].clone(), det.boxes[:n].clone(), det.scores[:n].clone(), det.labels[:n].clone()


RINGS = ("E0", "B0", "Y1E", "Y2M", "B1", "B2", "win_x", "win_boxes", "win_cnt", "glob_x")


@pytest.mark.parametrize("world,precision", [(2, "tf32"), (3, "tf32"), (2, "f16")])
def test_wavefront_engine_step_equals_sequential_step(world, precision):
    from mega_core.b200 import parallel, synth
    sd = synth.make_state_dict("mega_r101_tiny", seed=3)
    frames = 2 * world if world == 3 else 3 * world          # > memory_size = 4: the memory ring wraps
    with cpu_ops():
        solo = _make(sd, precision)
        _prime(solo, 1)
        payloads = [_payload(solo, 100 + t) for t in range(frames)]
        seq = []
        for t in range(frames):
            det = solo.dist_step(None, W_IMG, H_IMG, rank=0, world=1, payloads=payloads[t][None])[0]
            seq.append(_snap(solo, det))
        ranks = [_make(sd, precision) for _ in range(world)]     # "f16": the increment rows are fp16 inside 32-bit words
        for e in ranks:
            _prime(e, 1)
        # the memory rings live behind the per-frame scratch rows of these buffers
        offs = {"E0": solo.KP + solo.nl0, "B0": solo.KP + solo.nl0, "Y1E": solo.nq, "Y2M": solo.nq, "B1": solo.nl12,
                "B2": solo.nl12}
        for t0 in range(0, frames, world):
            gens = [ranks[r]._wave(None, W_IMG, H_IMG, r, world, payload=payloads[t0 + r]) for r in range(world)]
            dets = parallel.play(gens)
            for r in range(world):
                for a, b in zip(seq[t0 + r], _snap(ranks[r], dets[r])):
                    assert torch.equal(a, b), "key frame %d: wavefront differs from the sequential step" % (t0 + r)
            for r in range(1, world):                       # replicas agree after every group (ring part of the buffers)
                for name in RINGS:
                    o = offs.get(name, 0)
                    assert torch.equal(getattr(ranks[r], name)[o:], getattr(ranks[0], name)[o:]), (name, r, t0)
        assert seq[-1][1].shape[0] > 0, "degenerate test: no detections"
        for name in RINGS:                                  # and end in the sequential state
            o = offs.get(name, 0)
            assert torch.equal(getattr(ranks[0], name)[o:], getattr(solo, name)[o:]), name
        assert ranks[0].mem_pushed == solo.mem_pushed and list(ranks[0].win_slots) == list(solo.win_slots)


# ------------------------------------------------------------- the same through torch.distributed (gloo, 2 processes)
def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from cpu_ops import cpu_ops as ctx
    from mega_core.b200 import synth
    sd = synth.make_state_dict("mega_r101_tiny", seed=3)
    frames = 2 * world
    ok = True
    with ctx():
        solo = _make(sd)
        _prime(solo, 1)
        payloads = [_payload(solo, 100 + t) for t in range(frames)]
        seq = []
        for t in range(frames):
            seq.append(_snap(solo, solo.dist_step(None, W_IMG, H_IMG, rank=0, world=1, payloads=payloads[t][None])[0]))
        eng = _make(sd)
        _prime(eng, 1)
        # MegaEngine.dist_step_wave == parallel.drive(self._wave(...)): the per-frame branch is replaced by handing the
        # rank its payload (the backbone is not what this test is about), the collectives are real gloo all-gathers
        from mega_core.b200 import parallel
        for t0 in range(0, frames, world):
            det = parallel.drive(eng._wave(None, W_IMG, H_IMG, rank, world, payload=payloads[t0 + rank]))
            for a, b in zip(seq[t0 + rank], _snap(eng, det)):
                ok = ok and torch.equal(a, b)
        ok = ok and eng.mem_pushed == solo.mem_pushed
    q.put((rank, ok))
    dist.destroy_process_group()


def test_wavefront_engine_step_over_gloo():
    """world_size-2 gloo run of the wavefront step (parallel.drive + gather_payloads as NCCL would be driven)"""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_replicated_state_step_does_not_depend_on_world_size():
    """the DEFAULT multi-GPU step (dist_step: owner rows on the frame's owner, memory-feeding rows replicated) on the
    stand-ins: both ranks of a 2-rank group, played one after the other with the payloads handed over, give the
    1-rank results bit for bit (CPU counterpart of the GPU test of the same name in tests/test_engine_gpu.py)"""
    from mega_core.b200 import synth
    sd = synth.make_state_dict("mega_r101_tiny", seed=3)
    frames = 6
    with cpu_ops():
        solo = _make(sd)
        _prime(solo, 1)
        payloads = [_payload(solo, 100 + t) for t in range(frames)]
        seq, seq_k = [], []
        for t in range(frames):
            seq.append(_snap(solo, solo.dist_step(None, W_IMG, H_IMG, rank=0, world=1, payloads=payloads[t][None])[0]))
            seq_k.append(int(solo.cur_cnt.view(-1)[0]))         # live proposals of key frame t
        for rank in (0, 1):
            e = _make(sd)
            _prime(e, 1)
            for t in range(0, frames, 2):
                dets = e.dist_step(None, W_IMG, H_IMG, rank=rank, world=2, payloads=torch.stack(payloads[t:t + 2]))
                assert dets[1 - rank] is None
                # (rank 0 has meanwhile assembled the window of the foreign frame t + 1: its cur_cnt is that frame's)
                for a, b in zip(seq[t + rank], _snap(e, dets[rank], seq_k[t + rank])):
                    assert torch.equal(a, b), (rank, t)


def test_wave_selfcheck_function():
    """parallel.wave_selfcheck (what bench.py runs on every rank before it switches a multi-GPU run to the wavefront
    schedule) on the stand-ins: passes on the real engine, and reports a deliberately broken schedule"""
    from mega_core.b200 import parallel, synth
    sd = synth.make_state_dict("mega_r101_tiny", seed=3)
    with cpu_ops():
        ok, msg = parallel.wave_selfcheck(lambda: _make(sd), W_IMG, H_IMG, world=2, groups=3, use_graph=False)
        assert ok, msg
        orig = parallel.wave_tables

        def broken(*a, **k):                 # apply every increment BEFORE the reads: later frames' slots get clobbered
            t = orig(*a, **k)
            for key in ("0", "12", "b12"):
                both = torch.from_numpy(t["pre" + key]).maximum(torch.from_numpy(t["post" + key])).numpy()
                t["pre" + key], t["post" + key] = both, 0 * both - 1
            return t
        parallel.wave_tables = broken
        try:
            ok, msg = parallel.wave_selfcheck(lambda: _make(sd), W_IMG, H_IMG, world=2, groups=3, use_graph=False)
        finally:
            parallel.wave_tables = orig
        assert not ok and "differs" in msg


def test_switching_from_the_replicated_step_to_the_wavefront_step():
    """what bench.py does at N > 1: the video is primed with the replicated-state step (dist_step) and continues with the
    wavefront step -- both leave every rank in the sequential state, so they can be mixed"""
    from mega_core.b200 import parallel, synth
    sd = synth.make_state_dict("mega_r101_tiny", seed=3)
    world, frames = 2, 8
    with cpu_ops():
        solo = _make(sd)
        _prime(solo, 1)
        payloads = [_payload(solo, 100 + t) for t in range(frames)]
        seq, seq_k = [], []
        for t in range(frames):
            seq.append(_snap(solo, solo.dist_step(None, W_IMG, H_IMG, rank=0, world=1, payloads=payloads[t][None])[0]))
            seq_k.append(int(solo.cur_cnt.view(-1)[0]))
        ranks = [_make(sd) for _ in range(world)]
        for e in ranks:
            _prime(e, 1)
        for t0 in (0, 2):                                   # two groups with the replicated-state step
            for r in range(world):
                dets = ranks[r].dist_step(None, W_IMG, H_IMG, rank=r, world=world, payloads=torch.stack(payloads[t0:t0 + 2]))
                for a, b in zip(seq[t0 + r], _snap(ranks[r], dets[r], seq_k[t0 + r])):
                    assert torch.equal(a, b)
        for t0 in (4, 6):                                   # then the wavefront step
            dets = parallel.play([ranks[r]._wave(None, W_IMG, H_IMG, r, world, payload=payloads[t0 + r]) for r in range(world)])
            for r in range(world):
                for a, b in zip(seq[t0 + r], _snap(ranks[r], dets[r])):
                    assert torch.equal(a, b), (t0, r)
