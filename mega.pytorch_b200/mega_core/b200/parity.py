"""Replay of a committed reference fixture through a MegaEngine and the distance of its outputs from the reference's.

A fixture (tests/golden/mega_r101_*.pt, written by oracle/make_golden.py / oracle/make_golden_full.py from the UNMODIFIED
reference) holds, per key frame of a seeded synthetic video, the reference's detections and -- on its check frames --
the class logits / box deltas / proposals at the parity point of the north star (`FPNPredictor.forward`,
modeling/roi_heads/box_head/roi_box_predictors.py:50-57), in fp32 and (full-size fixture) also from the reference run in
fp64. Inputs are regenerated from the seeds, so the fixture is small. Used by tests/test_parity_full_gpu.py and by
bench.py's in-run `parity` block; nothing here touches oracle/.
"""
import torch

from . import synth


def match_rows(a, b, tol=0.75):
    """for each row of b (reference boxes) the index of a row of a within tol px (max over the 4 coordinates), or -1"""
    d = (a[:, None, :] - b[None, :, :]).abs().amax(2)
    val, idx = d.min(0)
    idx = idx.clone()
    idx[val > tol] = -1
    return idx


TOL = 1e-3          # the north star's tolerance on fp32 class logits


def _stats(d):
    """-> (max, 99th percentile, 99.9th percentile, fraction of entries beyond TOL)"""
    d = d.flatten().double()
    if d.numel() == 0:
        return 0.0, 0.0, 0.0, 0.0
    return (d.max().item(), torch.quantile(d, 0.99).item(), torch.quantile(d, 0.999).item(),
            (d > TOL).double().mean().item())


def replay(eng, gold, dev, frames=None, stop_after=None):
    """run the fixture's video through `eng` (a MegaEngine on `dev`) -> list of per-check-frame dicts:
    matched_frac (reference proposals reproduced within 0.75 px), logits_max / _p99 / _p999 / _frac_beyond_tol
    (|class logit - reference fp32| on matched rows; the last one = share of logits further than 1e-3), logits64_max / logits64_p99 (same against the reference run in fp64, when the fixture has it),
    deltas_max, dets / ref_dets, labels_equal"""
    h, w, total = gold["h"], gold["w"], gold["total"]
    gpf = gold["globals_per_frame"]
    if frames is None:
        frames = [synth.synthetic_frame(i, h, w).to(dev) for i in range(total)]
    n = len(gold["frames"]) if stop_after is None else min(stop_after, len(gold["frames"]))
    out = []
    for t in range(n):
        ref = gold["frames"][t]
        if t == 0:
            det = eng.start_video(frames[0], frames[1:13], [frames[j] for j in gpf[0]], w, h)
        else:
            det = eng.step(frames[min(t + 12, total - 1)], frames[gpf[t][0]], w, h)
        if "class_logits" not in ref:
            continue
        torch.cuda.synchronize(dev)
        k = int(eng.cur_cnt.view(-1)[0].item())
        props = eng.Bq0[:k].float().cpu()
        pred = eng.last_pred[:k].float().cpu()
        idx = match_rows(props, ref["proposals"])
        m = idx >= 0
        lmax, lp99, lp999, lfrac = _stats((pred[idx[m], :31] - ref["class_logits"][m]).abs())
        dmax = _stats((pred[idx[m], 31:155] - ref["box_regression"][m]).abs())[0]
        b, s, l = det.to_host()
        row = {"frame": t, "proposals": k, "ref_proposals": int(ref["proposals"].shape[0]),
               "matched_frac": m.float().mean().item(), "logits_max": lmax, "logits_p99": lp99, "logits_p999": lp999,
               "logits_frac_beyond_tol": lfrac, "deltas_max": dmax,
               "finite": bool(torch.isfinite(pred).all()), "dets": int(b.shape[0]), "ref_dets": int(ref["boxes"].shape[0]),
               "labels_equal": bool(b.shape[0] == ref["boxes"].shape[0] and torch.equal(l.cpu(), ref["labels"])),
               "logit_rms": ref["class_logits"].double().pow(2).mean().sqrt().item()}
        if "class_logits_fp64" in ref:
            i64 = match_rows(props, ref["proposals_fp64"])
            m64 = i64 >= 0
            (row["logits64_max"], row["logits64_p99"], row["logits64_p999"],
             row["logits64_frac_beyond_tol"]) = _stats((pred[i64[m64], :31] - ref["class_logits_fp64"][m64]).abs())
        out.append(row)
    return out


def summarize(rows):
    """worst case over the check frames"""
    if not rows:
        return None
    s = {"check_frames": len(rows), "min_matched_frac": min(r["matched_frac"] for r in rows),
         "logits_max": max(r["logits_max"] for r in rows), "logits_p99": max(r["logits_p99"] for r in rows),
         "logits_p999": max(r["logits_p999"] for r in rows),
         "logits_frac_beyond_tol": max(r["logits_frac_beyond_tol"] for r in rows),
         "deltas_max": max(r["deltas_max"] for r in rows), "all_finite": all(r["finite"] for r in rows),
         "frames_with_equal_labels": sum(r["labels_equal"] for r in rows),
         "frames_with_equal_det_count": sum(r["dets"] == r["ref_dets"] for r in rows)}
    if "logits64_max" in rows[0]:
        s["logits64_max"] = max(r["logits64_max"] for r in rows)
        s["logits64_p99"] = max(r["logits64_p99"] for r in rows)
        s["logits64_p999"] = max(r["logits64_p999"] for r in rows)
        s["logits64_frac_beyond_tol"] = max(r["logits64_frac_beyond_tol"] for r in rows)
    return s
