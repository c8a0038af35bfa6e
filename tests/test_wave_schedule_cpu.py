"""Host logic of the wavefront multi-GPU schedule (mega_core/b200/parallel.py: wave_tables / play / drive; SURVEY.md
section 8e option ii), checked on the CPU by symbolic execution: the memory rings hold frame numbers instead of
features, a stage's "output" is a hash of what it read, and the wavefront must reproduce sequential processing
-- at every read and in the final state -- for any world size and any fill level of the rings."""
import importlib.util
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "mega_parallel", os.path.join(ROOT, "mega.pytorch_b200", "mega_core", "b200", "parallel.py"))
par = importlib.util.module_from_spec(spec)
spec.loader.exec_module(par)

R0, R12, MEMF = 5, 2, 9          # small stand-ins for 75 / 15 rows per frame and 25 frames of memory
BASE0, BASE12, BASEB12 = 11, 9, 4


class Rings:
    """stage-0 / 1 / 2 memories as arrays of (frame, row, value-hash) codes"""

    def __init__(self):
        self.m = [np.zeros(BASE0 + MEMF * R0, np.int64), np.zeros(BASE12 + MEMF * R12, np.int64),
                  np.zeros(BASE12 + MEMF * R12, np.int64)]
        self.pushed = 0

    def read(self, stage, valid_frames):
        rows, base = (R0, BASE0) if stage == 0 else (R12, BASE12)
        return self.m[stage][base:base + valid_frames * rows].copy()

    def apply(self, stage, inc_all, table):
        for i, d in enumerate(table):
            if d >= 0:
                self.m[stage][d] = inc_all[i]


def H(*parts):
    return hash(tuple(p if isinstance(p, (int, str)) else tuple(np.asarray(p).tolist()) for p in parts)) % (1 << 40)


def rows_of(code, stage):
    return np.asarray([(code + 31 * j) % (1 << 40) for j in range(R0 if stage == 0 else R12)], dtype=np.int64)


def push(rings, stage, slot, inc):
    rows, base = (R0, BASE0) if stage == 0 else (R12, BASE12)
    rings.m[stage][base + slot * rows: base + (slot + 1) * rows] = inc


def sequential(rings, frames):
    """the reference's order per frame and stage: read the memory, then push (extractors :913-928): stage 0 pushes rows
    that need no memory (globally enhanced rows of the oldest local frame), stage s > 0 the output of stage s-1"""
    results = {}
    for t in frames:
        valid, slot = min(rings.pushed, MEMF), rings.pushed % MEMF
        out0 = H(t, 0, rings.read(0, valid))
        push(rings, 0, slot, rows_of(H(t, "inc0"), 0))
        out1 = H(t, 1, out0, rings.read(1, valid))
        push(rings, 1, slot, rows_of(out0, 1))
        out2 = H(t, 2, out1, rings.read(2, valid))
        push(rings, 2, slot, rows_of(out1, 2))
        results[t] = out2
        rings.pushed += 1
    return results


def wave_rank(rings, t0, rank, world):
    """generator: the schedule of MegaEngine._wave for one rank, on symbolic rings"""
    tabs = par.wave_tables(rings.pushed, rank, world, R0, R12, MEMF, BASE0, BASE12, BASE12)
    t = t0 + rank
    valid = int(tabs["valid"][rank])
    carry, later = None, []
    for s in range(3):
        key = "0" if s == 0 else "12"
        inc = rows_of(H(t, "inc0"), 0) if s == 0 else rows_of(carry, s)
        out = torch.zeros(world, inc.size, dtype=torch.int64)
        allinc = (yield torch.from_numpy(inc), out).reshape(-1).numpy().copy()
        rings.apply(s, allinc, tabs["pre" + key])                  # the group's earlier frames only
        view = rings.read(s, valid)
        carry = H(t, s, view) if s == 0 else H(t, s, carry, view)
        later.append((s, allinc, tabs["post" + key]))
    for s, allinc, post in later:                                   # own and later frames, after the last read
        rings.apply(s, allinc, post)
    rings.pushed += world
    return t, carry


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_wavefront_equals_sequential(world):
    n_groups = 6                       # crosses the point where the rings become full and wrap (MEMF = 9)
    seq = Rings()
    seq_results = sequential(seq, range(n_groups * world))
    ranks = [Rings() for _ in range(world)]
    for grp in range(n_groups):
        outs = par.play([wave_rank(ranks[r], grp * world, r, world) for r in range(world)])
        for r, (t, res) in enumerate(outs):
            assert t == grp * world + r and res == seq_results[t], (grp, r)
        for r in range(world):         # replicas agree after every group
            for s in range(3):
                assert np.array_equal(ranks[r].m[s], ranks[0].m[s])
    for s in range(3):
        assert np.array_equal(ranks[0].m[s], seq.m[s]), s
    assert ranks[0].pushed == seq.pushed


def test_wave_tables_reject_groups_larger_than_the_ring():
    with pytest.raises(AssertionError):
        par.wave_tables(0, 0, 10, R0, R12, MEMF, BASE0, BASE12, BASE12)     # two frames of a group would share a slot


def test_wave_tables_partition_the_group():
    tabs = par.wave_tables(23, 2, 4, 75, 15, 25, 2175, 675, 375)
    for key, rows in (("0", 75), ("12", 15), ("b12", 15)):
        pre, post = tabs["pre" + key], tabs["post" + key]
        assert ((pre >= 0) ^ (post >= 0)).all()                       # every increment row goes exactly one way
        assert (pre[:2 * rows] >= 0).all() and (pre[2 * rows:] < 0).all()
    assert tabs["pre0"][0] == 2175 + 23 * 75 and tabs["post0"][2 * 75] == 2175 + 0 * 75      # slot (23+2) % 25 = 0
    assert tabs["valid"].tolist() == [23, 24, 25, 25]
