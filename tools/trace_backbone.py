#!/usr/bin/env python
"""Per-layer timeline of the REAL backbone chain (res2-res4 + RPN head) of a MEGA R-101 step: the in-kernel event trace of
one CTA at layer granularity -> for every layer the time from its barrier to the next layer's barrier, its FLOPs and the
achieved rate, grouped by layer shape.   python tools/trace_backbone.py [--fps 1|2|4] [--cta 0]"""
import argparse
import collections
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mega_core._lib import lib  # noqa: E402
from mega_core.b200 import engine, ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fps", type=int, default=1)
ap.add_argument("--cta", type=int, default=0)
ap.add_argument("--layer", type=int, default=-1, help="print every event of this layer instead of the per-layer table")
args = ap.parse_args()
dev = torch.device("cuda:0")
h, w = 600, 1000
frames = [synth.synthetic_frame(i, h, w).to(dev) for i in range(16)]
pairs = [torch.cat([frames[(i + 12) % 16], frames[(5 * i + 3) % 16]], 0) for i in range(16)]
sd = synth.make_state_dict("mega_r101", seed=0)
eng = engine.MegaEngine(sd, engine.EngineConfig(precision="f16"), device=dev)
n = args.fps
with torch.no_grad():
    eng.start_video(frames[0], frames[1:13], [frames[(3 * j + 1) % 16] for j in range(10)], w, h)
    batch = torch.cat([pairs[j] for j in range(n)], 0)
    for _ in range(3):
        eng.stepn_batched(batch, w, h)
    torch.cuda.synchronize()
    # trace the backbone chain of the next step only
    # the chain that ran the batch of 2n images: one lane (95 layers) or two interleaved lanes of n images (190 layers)
    chains = [c for c in eng.backbone._chains.values()
              if c.n >= 90 and c.info["layers"][0]["m"] * (2 if c.depth == 2 and c.n >= 180 else 1) == 2 * n * 150 * 250]
    assert len(chains) == 1, [(c.n, c.depth, c.info["layers"][0]["m"]) for c in eng.backbone._chains.values()]
    ch = chains[0]
    trace = torch.zeros(3 * 4096 * 2, dtype=torch.int64, device=dev)
    lib.mega_conv_chain_set_trace2(ctypes.c_void_p(trace.data_ptr()), args.cta, 1 if args.layer < 0 else 2 + args.layer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ch.launch()
    e1.record()
    torch.cuda.synchronize()
    lib.mega_conv_chain_set_trace2(None, 0, 0)
t = trace.cpu().view(3, 4096, 2)
if args.layer >= 0:
    CODES = {1: "prod:layer_begin", 2: "prod:barrier_passed", 3: "prod:issue_kb", 7: "mma:kb_ready", 4: "epi:tile_acc_ready",
             5: "epi:residual_ready", 6: "epi:tile_done", 8: "epi:layer_tiles_done", 9: "epi:stores_drained", 10: "epi:arrived"}
    evs = []
    for role in range(3):
        for tag, clk in t[role].tolist():
            if tag or clk:
                evs.append((clk, role, (tag >> 8) & 0xffffff, tag & 0xff))
    evs.sort()
    print("layer", args.layer, ch.info["layers"][args.layer])
    t0, last = evs[0][0], evs[0][0]
    for clk, role, idx, code in evs:
        if code in (3, 7) and idx % 4:
            continue
        print("%9.3f us +%6.3f  %-22s idx %d" % ((clk - t0) / 1965.0, (clk - last) / 1965.0, CODES.get(code, code), idx))
        last = clk
    sys.exit(0)
ev = collections.defaultdict(dict)
for role in range(3):
    for tag, clk in t[role].tolist():
        if tag == 0 and clk == 0:
            continue
        ev[tag >> 32][(role, tag & 0xff)] = clk
layers = ch.info["layers"]
rows = []
keys = sorted(ev)
for i, l in enumerate(keys):
    e = ev[l]
    start = e.get((0, 2))
    nxt = ev[keys[i + 1]].get((0, 2)) if i + 1 < len(keys) else e.get((2, 10))
    if start is None or nxt is None:
        continue
    info = layers[l]
    gf = 2.0 * info["m"] * info["batch"] * info["cout"] * info["k"] * info["taps"] / 1e9
    us = (nxt - start) / 1965.0
    rows.append({"layer": l, "m": info["m"], "cout": info["cout"], "k": info["k"], "taps": info["taps"], "bn": info["bn"],
                 "sk": info["sk"], "res": info["res"], "us": round(us, 2), "gflop": round(gf, 3), "tflops": round(gf / us * 1e-3, 1),
                 "tiles_done_us": round((e.get((2, 8), start) - start) / 1965.0, 2),
                 "arrived_us": round((e.get((2, 10), start) - start) / 1965.0, 2)})
total = sum(r["us"] for r in rows)
print("chain %d layers, %.1f us (events), sum of traced layers %.1f us, grid %d" % (ch.n, e0.elapsed_time(e1) * 1e3, total, ch.grid))
grp = collections.OrderedDict()
for r in rows:
    k = (r["m"], r["cout"], r["k"], r["taps"], r["bn"], r["sk"], r["res"])
    g = grp.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += r["us"]; g[2] += r["gflop"]; g[3] += r["tiles_done_us"]; g[4] += r["arrived_us"]
print("%-52s %3s %9s %8s %8s %10s %10s" % ("(m, cout, k, taps, bn, sk, res)", "n", "us", "GF", "TF/s", "tiles_done", "arrived"))
for k, g in sorted(grp.items(), key=lambda kv: -kv[1][1]):
    print("%-52s %3d %9.1f %8.1f %8.1f %10.1f %10.1f" % (str(k), g[0], g[1], g[2], g[2] / g[1] * 1e3, g[3] / g[0], g[4] / g[0]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"chain_us": e0.elapsed_time(e1) * 1e3, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "backbone_trace_fps%d.json" % n), "w"))
