"""BASELINE-size parity fixture and the parity FLOOR of the reference itself.

Runs the UNMODIFIED reference (`/root/reference`, imported through oracle/ref_import.py) on the MEGA R-101
configuration BASELINE.json names -- 600x1000 synthetic video, 25 local / 10 global / 25 memory frames -- for
N key frames (default 42: the long-range memory is full from key frame 25 on), three ways:

  fp32      the reference as shipped (fp32, 8 threads)            -> tests/golden/mega_r101_600x1000.pt
  fp32_tN   the same with another thread count (summation order of the oneDNN / OpenMP reductions changes)
  fp64      the same modules after `model.double()` on double inputs. BoxList stores boxes as fp32
            (structures/bounding_box.py:22), so box coordinates are rounded to fp32 at every BoxList boundary
            exactly as in the fp32 run; three dtype casts are needed (rois -> the feature dtype for the compiled
            ROIAlign, dets / scores -> float for the compiled NMS, rois -> double at the entry of
            cal_position_embedding) and the default dtype is double; nothing else is touched.

The fixture stores, for every key frame, the detections (boxes, scores, labels) and for the CHECK frames also the
class logits / box deltas / proposals of the fp32 run AND of the fp64 run. `profiles/r02_parity_floor.json` holds
the floor: max / p99 |logit(fp32) - logit(fp64)| and |logit(fp32, 8 thr) - logit(fp32, N thr)| per check frame on
proposals matched by box -- the distance between two evaluations of the SAME reference that differ only in
rounding. A GPU arithmetic mode is at parity when its distance to the fp64 run is of that size.

Usage (here only; needs /root/reference):
  python oracle/make_golden_full.py run fp32 [--h 600 --w 1000 --frames 42 --threads 8]
  python oracle/make_golden_full.py run fp64
  python oracle/make_golden_full.py run fp32 --threads 3 --tag fp32_t3
  python oracle/make_golden_full.py assemble          # fixture + floor json from the runs under oracle/_full/
"""
import argparse
import json
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
WORK = os.path.join(HERE, "_full")          # scratch (git-ignored): one .pt per run

import ref_import  # noqa: E402
from make_golden import synth, _FakeImage  # noqa: E402

TOTAL = 64          # frames in the synthetic video (seg_len)


def check_frames(n):
    base = [0, 1, 2, 12, 13, 24, 25, 26, 27, 33, 38, 39, 40, 41]
    return [t for t in base if t < n] + ([n - 1] if n - 1 not in base else [])


def run(tag, dtype, h, w, n_frames, threads):
    torch.set_num_threads(threads)
    cfg = ref_import.build_cfg("configs/MEGA/vid_R_101_C4_MEGA_1x.yaml")
    from mega_core.modeling.detector import build_detection_model
    from mega_core.structures.image_list import to_image_list
    import mega_core.modeling.detector.generalized_rcnn_mega as gm
    model = build_detection_model(cfg).eval()
    sd = synth.make_state_dict("mega_r101", seed=0)
    full = dict(sd)
    full["rpn.anchor_generator.cell_anchors.0"] = model.state_dict()["rpn.anchor_generator.cell_anchors.0"]
    model.load_state_dict(full, strict=True)
    if dtype == torch.float64:
        # (torch.set_default_dtype(float64) is NOT used: it changes the proposals of the reference)
        model = model.double()
        import mega_core.modeling.roi_heads.box_head.roi_box_feature_extractors as fx
        orig_pe = fx.AttentionExtractor.cal_position_embedding
        fx.AttentionExtractor.cal_position_embedding = lambda self, r1, r2: orig_pe(self, r1.double(), r2.double())
        import mega_core.layers  # noqa: F401
        import mega_core.structures.boxlist_ops  # noqa: F401
        ra = sys.modules["mega_core.layers.roi_align"]      # (the package attribute of that name is the function)
        bo = sys.modules["mega_core.structures.boxlist_ops"]
        _C = sys.modules["mega_core._C"]
        orig_fwd = ra._ROIAlign.forward

        def fwd(ctx, input, roi, *a):       # rois follow the feature dtype (ROIAlign_cpu.cpp reads both as T)
            return orig_fwd(ctx, input, roi.to(input.dtype), *a)

        ra._ROIAlign.forward = staticmethod(fwd)
        orig_nms = bo._box_nms
        bo._box_nms = lambda b, s, t: orig_nms(b.float(), s.float(), t)
    frames = [synth.synthetic_frame(i, h, w).to(dtype) for i in range(TOTAL)]
    gidx = synth.global_frame_indices(TOTAL, seed=0)
    gpf = [gidx[:10]] + [[gidx[(10 + t - 1) % TOTAL]] for t in range(1, n_frames)]
    gm.Image = types.SimpleNamespace(open=lambda path: _FakeImage(int(os.path.basename(path).split(".")[0])))
    hooks = {}
    pred = model.roi_heads.box.predictor
    orig_pred = pred.forward

    def pred_fwd(x):
        r = orig_pred(x)
        hooks["class_logits"], hooks["box_regression"] = r[0].clone(), r[1].clone()
        return r

    pred.forward = pred_fwd
    pp = model.roi_heads.box.post_processor
    orig_pp = pp.forward

    def pp_fwd(x, boxes):
        hooks["proposals"] = boxes[0].bbox.clone()
        return orig_pp(x, boxes)

    pp.forward = pp_fwd
    outs, secs = [], []
    os.makedirs(WORK, exist_ok=True)
    with torch.no_grad():
        for t in range(n_frames):
            t0 = time.time()
            images = {"cur": frames[t][0].clone(),
                      "ref_l": [] if t == 0 else [to_image_list(frames[min(t + 12, TOTAL - 1)][0].clone())],
                      "ref_g": [to_image_list(frames[j][0].clone()) for j in gpf[t]],
                      "frame_category": 0 if t == 0 else 1, "seg_len": TOTAL, "pattern": "%06d",
                      "img_dir": "/nonexistent/%s.JPEG", "transforms": lambda im: frames[im.idx][0].clone()}
            res = model(images)[0]
            secs.append(time.time() - t0)
            outs.append({"boxes": res.bbox.clone(), "scores": res.get_field("scores").clone(),
                         "labels": res.get_field("labels").clone(), **{k: v.clone() for k, v in hooks.items()}})
            print("[%s] frame %d: %.1f s, %d proposals, %d detections" % (tag, t, secs[-1], hooks["proposals"].shape[0],
                                                                        res.bbox.shape[0]), flush=True)
            if t % 8 == 7 or t == n_frames - 1:
                torch.save({"tag": tag, "h": h, "w": w, "threads": threads, "globals_per_frame": gpf, "frames": outs,
                            "secs": secs}, os.path.join(WORK, "%s_%dx%d.pt" % (tag, h, w)))


def _match(a, b, tol=0.75):
    """for each row of b the index of a row of a within tol px (max-abs over the 4 coordinates), or -1"""
    d = (a.double()[:, None, :] - b.double()[None, :, :]).abs().amax(2)
    val, idx = d.min(0)
    idx[val > tol] = -1
    return idx


def _dist(x, y):
    d = (x.double() - y.double()).abs().flatten()
    if d.numel() == 0:
        return {"max": 0.0, "p99": 0.0, "p999": 0.0, "frac_beyond_1e-3": 0.0, "mean": 0.0}
    return {"max": d.max().item(), "p99": torch.quantile(d, 0.99).item(), "p999": torch.quantile(d, 0.999).item(),
            "frac_beyond_1e-3": (d > 1e-3).double().mean().item(), "mean": d.mean().item()}


def assemble(h, w):
    base = torch.load(os.path.join(WORK, "fp32_%dx%d.pt" % (h, w)))
    others = {}
    for f in sorted(os.listdir(WORK)):
        if f.endswith("_%dx%d.pt" % (h, w)) and not f.startswith("fp32_%d" % h):
            o = torch.load(os.path.join(WORK, f))
            others[o["tag"]] = o
    n = len(base["frames"])
    chk = check_frames(n)
    floor = {"what": "distance between evaluations of the UNMODIFIED reference (MEGA R-101, %dx%d, synthetic video, "
                     "seed 0) that differ only in rounding; class logits on proposals matched by box (0.75 px)" % (h, w),
             "frames": n, "check_frames": chk, "base": {"tag": "fp32", "threads": base["threads"]}, "runs": {}}
    for tag, o in others.items():
        m = min(n, len(o["frames"]))
        rows = []
        for t in range(m):
            a, b = base["frames"][t], o["frames"][t]
            idx = _match(b["proposals"], a["proposals"])        # for every base proposal, its row in the other run
            ok = idx >= 0
            dl = _dist(b["class_logits"][idx[ok]], a["class_logits"][ok])
            db = _dist(b["box_regression"][idx[ok]], a["box_regression"][ok])
            rows.append({"frame": t, "proposals": int(a["proposals"].shape[0]), "other_proposals": int(b["proposals"].shape[0]),
                         "matched_frac": ok.float().mean().item(), "logits": dl, "deltas": db,
                         "dets": int(a["boxes"].shape[0]), "other_dets": int(b["boxes"].shape[0]),
                         "labels_equal": bool(a["labels"].shape == b["labels"].shape and torch.equal(a["labels"], b["labels"])),
                         "logit_rms": a["class_logits"].double().pow(2).mean().sqrt().item()})
        floor["runs"][tag] = {"threads": o["threads"], "frames": rows,
                              "logits_max_over_frames": max(r["logits"]["max"] for r in rows),
                              "logits_p99_max_over_frames": max(r["logits"]["p99"] for r in rows),
                              "logits_p999_max_over_frames": max(r["logits"]["p999"] for r in rows),
                              "logits_frac_beyond_1e-3_max_over_frames": max(r["logits"]["frac_beyond_1e-3"] for r in rows),
                              "frames_with_a_logit_beyond_1e-3": sum(r["logits"]["max"] > 1e-3 for r in rows),
                              "min_matched_frac": min(r["matched_frac"] for r in rows),
                              "frames_with_equal_labels": sum(r["labels_equal"] for r in rows)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    name = "r02_parity_floor.json" if (h, w) == (600, 1000) else "r02_parity_floor_%dx%d.json" % (h, w)
    with open(os.path.join(ROOT, "profiles", name), "w") as fh:
        json.dump(floor, fh, indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "frames"} for k, v in floor["runs"].items()}, indent=1))
    f64 = others.get("fp64")
    gold = {"arch": "mega_r101", "seed": 0, "h": h, "w": w, "total": TOTAL, "globals_per_frame": base["globals_per_frame"],
            "check_frames": chk, "frames": []}
    for t in range(n):
        a = base["frames"][t]
        e = {"boxes": a["boxes"], "scores": a["scores"], "labels": a["labels"], "n_proposals": int(a["proposals"].shape[0])}
        if t in chk:
            e.update({"class_logits": a["class_logits"], "box_regression": a["box_regression"], "proposals": a["proposals"]})
            if f64 is not None and t < len(f64["frames"]):
                b = f64["frames"][t]
                e.update({"class_logits_fp64": b["class_logits"].float(), "proposals_fp64": b["proposals"].float()})
        gold["frames"].append(e)
    out = os.path.join(GOLD, "mega_r101_%dx%d.pt" % (h, w))
    torch.save(gold, out)
    print("wrote", out, os.path.getsize(out), "bytes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["run", "assemble"])
    ap.add_argument("mode", nargs="?", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--h", type=int, default=600)
    ap.add_argument("--w", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=42)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--tag", default=None)
    a = ap.parse_args()
    if a.cmd == "assemble":
        return assemble(a.h, a.w)
    run(a.tag or a.mode, torch.float64 if a.mode == "fp64" else torch.float32, a.h, a.w, a.frames, a.threads)


if __name__ == "__main__":
    main()
