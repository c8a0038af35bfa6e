#!/usr/bin/env python
"""In-kernel timeline of the persistent chain kernel (diagnostics): runs a few res4-shaped bottleneck blocks as one
chain with the event trace of one CTA switched on, and prints / saves per-event SM-clock timestamps.
    python tools/trace_chain.py [--cta 0] [--blocks 4] -> gpurun_out/chain_trace.json"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mega_core._lib import lib  # noqa: E402
from mega_core.b200 import ops  # noqa: E402

CODES = {1: "prod:layer_begin", 2: "prod:barrier_passed", 3: "prod:issue_kb", 7: "mma:kb_ready", 4: "epi:tile_acc_ready",
         5: "epi:residual_ready", 6: "epi:tile_done", 8: "epi:layer_tiles_done", 9: "epi:stores_drained",
         10: "epi:arrived"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cta", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--bn1", type=int, default=128)
    ap.add_argument("--bn2", type=int, default=128)
    ap.add_argument("--bn3", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n, h, w, c, mid = 2, 38, 63, 1024, 256
    x0 = torch.randn(n, h, w, c, generator=g).half().to(dev)
    ws, bufs = [], []
    for b in range(args.blocks):
        ws.append(((torch.randn(1, mid, c, generator=g) / c ** 0.5).half().to(dev),
                   (torch.randn(9, mid, mid, generator=g) / (9 * mid) ** 0.5).half().to(dev),
                   (torch.randn(1, c, mid, generator=g) / mid ** 0.5).half().to(dev)))
        bufs.append((torch.zeros(n, h, w, mid, device=dev, dtype=torch.float16),
                     torch.zeros(n, h, w, mid, device=dev, dtype=torch.float16),
                     torch.zeros(n, h, w, c, device=dev, dtype=torch.float16)))
    sc = torch.ones(c, device=dev)

    def run():
        x = x0
        for (w1, w2, w3), (t1, t2, y) in zip(ws, bufs):
            ops.conv_gemm(x, w1, t1, scale=sc[:mid], bias=sc[:mid], relu=True, block_n=args.bn1, stream_k=0)
            ops.conv_gemm(t1, w2, t2, taps=(3, 3), pad=1, scale=sc[:mid], bias=sc[:mid], relu=True, block_n=args.bn2, stream_k=0)
            ops.conv_gemm(t2, w3, y, scale=sc, bias=sc, residual=x, relu=True, block_n=args.bn3, stream_k=0)
            x = y

    cache = {}
    for _ in range(3):     # record + warm replays
        with ops.chain(cache, "k", dev):
            run()
    torch.cuda.synchronize()
    trace = torch.zeros(3 * 4096 * 2, dtype=torch.int64, device=dev)
    lib.mega_conv_chain_set_trace(ctypes.c_void_p(trace.data_ptr()), args.cta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with ops.chain(cache, "k", dev):
        run()
    e1.record()
    torch.cuda.synchronize()
    lib.mega_conv_chain_set_trace(None, 0)
    t = trace.cpu().view(3, 4096, 2)
    ev = []
    for role in range(3):
        for tag, clk in t[role].tolist():
            if tag == 0 and clk == 0:
                continue
            ev.append({"role": role, "layer": tag >> 32, "idx": (tag >> 8) & 0xffffff, "code": tag & 0xff, "clk": clk})
    t0 = min(e["clk"] for e in ev)
    for e in ev:
        e["us"] = round((e["clk"] - t0) / 1965.0, 3)     # SM clock 1965 MHz under load
        e["what"] = CODES.get(e["code"], "?")
    ev.sort(key=lambda e: e["clk"])
    out = {"chain_ms": e0.elapsed_time(e1), "layers": cache["k"].n, "grid": cache["k"].grid, "cta": args.cta, "events": ev}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "chain_trace_cta%d.json" % args.cta), "w") as fh:
        json.dump(out, fh)
    print("chain of %d layers: %.1f us total (%.1f us / layer), grid %d" % (cache["k"].n, out["chain_ms"] * 1e3,
                                                                         out["chain_ms"] * 1e3 / cache["k"].n, cache["k"].grid))
    last = None
    for e in ev:
        if e["layer"] in (3, 4, 5):     # one steady-state block
            print("%8.3f us  +%6.3f  L%d %-22s idx %d" % (e["us"], e["us"] - (last or e["us"]), e["layer"], e["what"], e["idx"]))
            last = e["us"]


if __name__ == "__main__":
    main()
