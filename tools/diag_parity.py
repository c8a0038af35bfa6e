"""GPU diagnostic: engine (TF32 kernels) and engine with the fp32 GEMM shadow vs the reference fixture /
the oracle, at several probe points. Writes gpurun_out/diag_parity.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mega.pytorch_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mega_oracle as mo  # noqa: E402
from fp32_shadow import fp32_shadow  # noqa: E402
from mega_core.b200 import engine, synth  # noqa: E402

dev = torch.device("cuda:0")
gold = torch.load(os.path.join(ROOT, "tests", "golden", "mega_r101_192x320.pt"))
h, w, total = gold["h"], gold["w"], gold["total"]
sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
dframes = [f.to(dev) for f in frames]
gpf = gold["globals_per_frame"]
orc = mo.MegaOracle(sd, mo.Cfg(cuda_nms_semantics=True), record=True)
traces = []
for t in range(3):
    infos = {"frame_category": 0 if t == 0 else 1, "ref_l": frames[1:13] if t == 0 else [frames[min(t + 12, total - 1)]],
             "ref_g": [frames[j] for j in gpf[t]]}
    orc.forward(frames[t], infos)
    traces.append({k: v.clone() for k, v in orc.trace.items()})


def match(a, b, tol=0.75):
    d = (a[:, None, :] - b[None, :, :]).abs().amax(2)
    val, idx = d.min(0)
    idx[val > tol] = -1
    return idx


def run(label, ctx):
    out = []
    with ctx:
        eng = engine.MegaEngine(sd, device=dev)
        for t in range(3):
            if t == 0:
                eng.start_video(dframes[0], dframes[1:13], [dframes[j] for j in gpf[0]], w, h)
            else:
                eng.step(dframes[min(t + 12, total - 1)], dframes[gpf[t][0]], w, h)
            torch.cuda.synchronize()
            tr = traces[t]
            k = int(eng.cur_cnt.view(-1)[0].item())
            props = eng.Bq0[:k].cpu()
            idx = match(props, tr["proposals"])
            m = idx >= 0
            kslot = list(eng.win_slots)[eng.cfg.key_frame_location]
            xfc = eng.win_x[kslot * eng.KP: kslot * eng.KP + k].cpu()
            pred = eng.last_pred[:k].cpu()
            x4 = eng.X4[:k].cpu()
            rec = {"frame": t, "k": k, "k_ref": int(tr["proposals"].shape[0]), "matched": m.float().mean().item(),
                   "prop_maxdiff_matched": (props[idx[m]] - tr["proposals"][m]).abs().max().item(),
                   "x_key_fc_err": (xfc[idx[m]] - tr["x_key_fc"][m]).abs().max().item(),
                   "x_key_fc_rms": tr["x_key_fc"].pow(2).mean().sqrt().item(),
                   "x_final_err": (x4[idx[m]] - tr["x_final"][m]).abs().max().item(),
                   "x_final_rms": tr["x_final"].pow(2).mean().sqrt().item(),
                   "logits_err": (pred[idx[m], :31] - tr["class_logits"][m]).abs().max().item(),
                   "logits_rms": tr["class_logits"].pow(2).mean().sqrt().item()}
            out.append(rec)
            print(label, rec)
    return out


import contextlib  # noqa: E402
res = {"tf32": run("tf32", contextlib.nullcontext()), "fp32_shadow": run("fp32", fp32_shadow())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "diag_parity.json"), "w"), indent=1)
