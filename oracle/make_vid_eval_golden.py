"""Runs the UNMODIFIED reference evaluator (mega_core/data/datasets/evaluation/vid/vid_eval.py) on seeded synthetic
detections and writes inputs + outputs to tests/golden/vid_eval.pt -- the fixture that pins
mega.pytorch_b200/mega_core/data/datasets/evaluation/vid/vid_eval.py. TEST INFRASTRUCTURE.
Usage: python oracle/make_vid_eval_golden.py   (needs /root/reference; run in a fresh process: it imports the reference's
`mega_core`, which cannot coexist with the product's package of the same name)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def synth_case(seed, n_images=60, n_classes=6, w=640, h=360):
    g = np.random.default_rng(seed)
    images = []
    for _ in range(n_images):
        n_gt = int(g.integers(0, 5))
        x1 = g.uniform(0, w - 80, n_gt)
        y1 = g.uniform(0, h - 80, n_gt)
        gt = np.stack([x1, y1, x1 + g.uniform(20, 200, n_gt), y1 + g.uniform(20, 150, n_gt)], 1).round().astype(np.float32) \
            if n_gt else np.zeros((0, 4), np.float32)
        gl = g.integers(1, n_classes, n_gt).astype(np.int64)
        motion = g.choice([0.3, 0.65, 0.7, 0.8, 0.9, 0.95, 1.0], n_gt).tolist()
        boxes, labels, scores = [], [], []
        for k in range(n_gt):                                 # detections around every object, some with the wrong label
            for _ in range(int(g.integers(0, 5))):
                boxes.append(gt[k] + g.normal(0, 6, 4).astype(np.float32))
                labels.append(gl[k] if g.random() < 0.8 else g.integers(1, n_classes))
                scores.append(np.float32(np.round(g.uniform(0.05, 1.0), 2)))     # 2 decimals: score ties do occur
        for _ in range(int(g.integers(0, 12))):               # clutter
            a, b = g.uniform(0, w - 60), g.uniform(0, h - 60)
            boxes.append(np.asarray([a, b, a + g.uniform(10, 150), b + g.uniform(10, 120)], np.float32))
            labels.append(g.integers(1, n_classes))
            scores.append(np.float32(np.round(g.uniform(0.0, 0.6), 2)))
        images.append({"size": (w, h), "gt": torch.from_numpy(gt), "gt_labels": torch.from_numpy(gl), "motion": motion,
                       "boxes": torch.from_numpy(np.asarray(boxes, np.float32).reshape(-1, 4)),
                       "labels": torch.from_numpy(np.asarray(labels, np.int64)),
                       "scores": torch.from_numpy(np.asarray(scores, np.float32))})
    return images


def main():
    ref_import.setup()
    # the file itself, by path: its package's __init__ chain pulls in COCO / Cityscapes helpers that are not installed
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_vid_eval", os.path.join(ref_import.REFERENCE, "mega_core", "data", "datasets", "evaluation", "vid", "vid_eval.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from mega_core.structures.bounding_box import BoxList
    cases = []
    for seed in (1, 2):
        images = synth_case(seed)
        gts, preds = [], []
        for im in images:
            gt = BoxList(im["gt"], im["size"], mode="xyxy")
            gt.add_field("labels", im["gt_labels"])
            pr = BoxList(im["boxes"], im["size"], mode="xyxy")
            pr.add_field("labels", im["labels"])
            pr.add_field("scores", im["scores"])
            gts.append(gt)
            preds.append(pr)
        out = {}
        with contextlib.redirect_stdout(io.StringIO()):
            prec, rec = ref.calc_detection_vid_prec_rec(gts, preds, None, 0.5, [0.0, 1.0])
            out["all"] = {"prec": prec, "rec": rec, "ap": ref.calc_detection_vid_ap(prec, rec),
                          "ap07": ref.calc_detection_vid_ap(prec, rec, use_07_metric=True)}
            motion = [im["motion"] for im in images]
            for name, rng in (("fast", [0.0, 0.7]), ("medium", [0.7, 0.9]), ("slow", [0.9, 1.0])):
                prec, rec = ref.calc_detection_vid_prec_rec(gts, preds, motion, 0.5, rng)
                out[name] = {"prec": prec, "rec": rec, "ap": ref.calc_detection_vid_ap(prec, rec)}
        cases.append({"seed": seed, "images": images, "reference": out})
        print("seed", seed, "mAP(all) %.4f" % np.nanmean(out["all"]["ap"]),
              " ".join("%s %.4f" % (k, np.nanmean(out[k]["ap"])) for k in ("fast", "medium", "slow")))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "vid_eval.pt")
    torch.save(cases, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
