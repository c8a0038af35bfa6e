"""mega_core.data.datasets.evaluation.vid (native matching in libmega_b200.so + array bookkeeping; SURVEY.md section 8f
row 2) against the outputs of the UNMODIFIED reference evaluator on the same seeded synthetic detections
(tests/golden/vid_eval.pt, written by oracle/make_vid_eval_golden.py): precision / recall arrays and APs must be equal
element for element, for the plain protocol and for the three motion-IoU ranges (ignore flags, fractional weights)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _boxlists(images):
    from mega_core.structures.bounding_box import BoxList
    gts, preds = [], []
    for im in images:
        gt = BoxList(im["gt"], im["size"], mode="xyxy")
        gt.add_field("labels", im["gt_labels"])
        pr = BoxList(im["boxes"], im["size"], mode="xyxy")
        pr.add_field("labels", im["labels"])
        pr.add_field("scores", im["scores"])
        gts.append(gt)
        preds.append(pr)
    return gts, preds


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True)


def test_vid_evaluator_equals_reference():
    from mega_core.data.datasets.evaluation.vid import calc_detection_vid_ap, calc_detection_vid_prec_rec
    cases = torch.load(os.path.join(ROOT, "tests", "golden", "vid_eval.pt"), weights_only=False)
    for case in cases:
        gts, preds = _boxlists(case["images"])
        ref = case["reference"]
        prec, rec = calc_detection_vid_prec_rec(gts, preds, None, 0.5, [0.0, 1.0])
        _same(prec, ref["all"]["prec"])
        _same(rec, ref["all"]["rec"])
        assert np.array_equal(calc_detection_vid_ap(prec, rec), ref["all"]["ap"], equal_nan=True)
        assert np.allclose(calc_detection_vid_ap(prec, rec, use_07_metric=True), ref["all"]["ap07"], equal_nan=True, atol=1e-12)
        motion = [im["motion"] for im in case["images"]]
        for name, rng in (("fast", [0.0, 0.7]), ("medium", [0.7, 0.9]), ("slow", [0.9, 1.0])):
            prec, rec = calc_detection_vid_prec_rec(gts, preds, motion, 0.5, rng)
            _same(prec, ref[name]["prec"])
            _same(rec, ref[name]["rec"])
            assert np.array_equal(calc_detection_vid_ap(prec, rec), ref[name]["ap"], equal_nan=True)


def test_eval_detection_vid_entry_point():
    from mega_core.data.datasets.evaluation.vid import eval_detection_vid
    case = torch.load(os.path.join(ROOT, "tests", "golden", "vid_eval.pt"), weights_only=False)[0]
    gts, preds = _boxlists(case["images"])
    res = eval_detection_vid(preds, gts, 0.5, [[0.0, 1.0]], motion_specific=False)
    assert abs(res[0]["map"] - np.nanmean(case["reference"]["all"]["ap"])) < 1e-15
    motion = [im["motion"] for im in case["images"]]
    res = eval_detection_vid(preds, gts, 0.5, [[0.0, 0.7], [0.9, 1.0]], motion_specific=True, motion_ious=motion)
    assert abs(res[0]["map"] - np.nanmean(case["reference"]["fast"]["ap"])) < 1e-15
    assert abs(res[1]["map"] - np.nanmean(case["reference"]["slow"]["ap"])) < 1e-15
