"""One launch of the backbone chain of a MEGA R-101 step under the profiler only (cudaProfilerStart/Stop around it):
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/chain python tools/ncu_chain.py --fps 4
--step: profile every kernel of one whole steady-state step instead (use with --metrics gpu__time_duration.sum)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mega_core.b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fps", type=int, default=4)
ap.add_argument("--step", action="store_true")
ap.add_argument("--precision", default="f16")
args = ap.parse_args()
dev = torch.device("cuda:0")
h, w = 600, 1000
frames = [synth.synthetic_frame(i, h, w).to(dev) for i in range(16)]
pairs = [torch.cat([frames[(i + 12) % 16], frames[(5 * i + 3) % 16]], 0) for i in range(16)]
sd = synth.make_state_dict("mega_r101", seed=0)
eng = engine.MegaEngine(sd, engine.EngineConfig(precision=args.precision), device=dev)
n = args.fps
with torch.no_grad():
    eng.start_video(frames[0], frames[1:13], [frames[(3 * j + 1) % 16] for j in range(10)], w, h)
    batch = torch.cat([pairs[j] for j in range(n)], 0)
    for _ in range(3):
        eng.stepn_batched(batch, w, h)
    torch.cuda.synchronize()
    if args.step:
        torch.cuda.profiler.start()
        eng.stepn_batched(batch, w, h)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        ch = max(eng.backbone._chains.values(), key=lambda c: c.flops)
        print("profiling chain: %d layers, depth %d, grid %d, %.1f GFLOP" % (ch.n, ch.depth, ch.grid, ch.flops / 1e9))
        torch.cuda.profiler.start()
        ch.launch()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
