"""strict mode (fp32x3): accuracy on the 600x1000 reference fixture and speed as a function of the accumulator segment
length (mega_set_split3_seg_len).   python tools/strict_probe.py [seg_len ...]  -> gpurun_out/strict_probe.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mega_core._lib import lib  # noqa: E402
from mega_core.b200 import engine, parity, synth  # noqa: E402

dev = torch.device("cuda:0")
gold = torch.load(os.path.join(ROOT, "tests", "golden", "mega_r101_600x1000.pt"))
frames = [synth.synthetic_frame(i, gold["h"], gold["w"]).to(dev) for i in range(gold["total"])]
sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
out = {}
for seg in [int(a) for a in sys.argv[1:]] or [4, 2, 1]:
    lib.mega_set_split3_seg_len(seg)
    eng = engine.MegaEngine(sd, engine.EngineConfig(precision="fp32x3"), device=dev)
    with torch.no_grad():
        rows = parity.replay(eng, gold, dev, frames=frames)
        s = parity.summarize(rows)
        eng.use_graph = True
        pairs = [torch.cat([frames[(i + 12) % 64], frames[(5 * i + 3) % 64]], 0) for i in range(8)]
        for i in range(4):
            eng.step_batched(pairs[i % 8], gold["w"], gold["h"])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            eng.step_batched(pairs[i % 8], gold["w"], gold["h"])
        e1.record()
        torch.cuda.synchronize()
    s["ms_per_step"] = e0.elapsed_time(e1) / 10
    out[seg] = {"summary": s, "frames": rows}
    print("seg_len", seg, json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in s.items()}), flush=True)
    del eng
    torch.cuda.empty_cache()
lib.mega_set_split3_seg_len(4)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "strict_probe.json"), "w"), indent=1)
