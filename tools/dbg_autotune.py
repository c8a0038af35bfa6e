import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mega.pytorch_b200")):
    sys.path.insert(0, p)
import torch
from mega_core.b200 import engine, ops, synth
ops.AUTOTUNE[0] = True
dev = torch.device("cuda:0")
sd = synth.make_state_dict("mega_r101_tiny", seed=2)
img = torch.cat([synth.synthetic_frame(1, 96, 160), synth.synthetic_frame(2, 96, 160)], 0).to(dev)
bb = engine.Backbone(sd, dev, dtype=torch.float16)
for i in range(2):
    y = bb.forward(img)
    torch.cuda.synchronize()
    print("pass", i, float(y.float().abs().mean()))
print("tuned entries", len(ops.TUNED))
