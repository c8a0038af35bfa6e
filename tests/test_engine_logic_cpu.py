"""Host logic of the engines on the CPU stand-ins (tests/cpu_ops.py) against the oracle / the reference's fixtures:
buffer layouts, weight packing, FlowNetS wiring (row-slab stem, strided convolutions, the four parity classes of every
transposed convolution, crops, concats), key-frame caching -- everything except the CUDA kernels themselves, whose
parity tests are the -m gpu suite. Contractions run through the exact-fp32 shadow, so the expected agreement with the
reference's outputs is the fp32 noise of a different summation order."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpu_ops import cpu_ops  # noqa: E402


def _match_rows(a, b, tol=0.75):
    d = (a[:, None, :] - b[None, :, :]).abs().amax(2)
    val, idx = d.min(0)
    idx[val > tol] = -1
    return idx


def test_dff_engine_logic_matches_reference_fixture():
    """DffEngine vs the unmodified reference's GeneralizedRCNNDFF outputs: key, non-key,
    non-key, key frames -- flow, scale map, warped features, every proposal, class logits within 1e-3, detections"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "dff_r101_192x320.pt"))
    h, w = gold["h"], gold["w"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    with cpu_ops():
        eng = engine.DffEngine(sd, engine.EngineConfig(precision="tf32"), device="cpu")
        for t, (key, ref) in enumerate(zip(gold["key_flags"], gold["frames"])):
            if t >= 4:
                break
            det = eng.forward(synth.synthetic_frame(gold["frame_stride"] * t, h, w), key, w, h)
            k = int(eng.last_cnt[0])
            idx = _match_rows(eng.last_props[:k], ref["proposals"])
            assert (idx >= 0).all() and k == ref["proposals"].shape[0], "frame %d: proposals differ" % t
            flow = eng.last_flow[..., :2].permute(0, 3, 1, 2)
            assert (flow - ref["flow"]).abs().max() < 1e-4
            assert (eng.last_scale.permute(0, 3, 1, 2)[:, ::64] - ref["scale_sample"]).abs().max() < 1e-4
            assert (eng.last_feats.permute(0, 3, 1, 2)[:, ::64] - ref["feats_sample"]).abs().max() < 1e-3 * max(ref["feats_rms"], 1)
            assert (eng.last_pred[:k][idx, :31] - ref["class_logits"]).abs().max() < 1e-3
            n = int(det.count[0])
            assert n == ref["boxes"].shape[0] and torch.equal(det.labels[:n], ref["labels"])
            assert torch.allclose(det.boxes[:n], ref["boxes"], atol=2e-2)


def _check_frame(eng, det, ref, t, logit_tol=1e-3):
    k = int(eng.cur_cnt.view(-1)[0]) if hasattr(eng, "cur_cnt") else int(eng.last_cnt[0])
    props = eng.Bq0[:k] if hasattr(eng, "Bq0") else eng.last_props[:k]
    idx = _match_rows(props, ref["proposals"])
    assert (idx >= 0).all() and k == ref["proposals"].shape[0], "frame %d: proposals differ" % t
    assert (eng.last_pred[:k][idx, :31] - ref["class_logits"]).abs().max() < logit_tol, "frame %d: class logits" % t
    n = int(det.count.reshape(-1)[0])
    assert n == ref["boxes"].shape[0] and torch.equal(det.labels[:n], ref["labels"]), "frame %d: detections" % t


def test_mega_module_and_engine_logic_match_reference_fixture():
    """The hot path through the reference-facing module API: GeneralizedRCNNMEGA built by
    build_detection_model(cfg), weights through load_state_dict, `model(images)` with the dict VIDMEGADataset builds
    (frame 0 with its look-ahead and global frames, then a steady-state frame) -> list[BoxList]; underneath, MegaEngine's
    start_video / step: index tables, window / global rings, long-range memory pushes, fused aggregation. Against the
    unmodified reference's GeneralizedRCNNMEGA outputs."""
    from mega_core.b200 import engine, synth
    from mega_core.modeling.detector import build_detection_model_from_state_dict
    from mega_core.modeling.nets import engine_config_from
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "mega_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    gpf = gold["globals_per_frame"]
    with cpu_ops():
        model = build_detection_model_from_state_dict(sd, method="mega", device="cpu", precision="tf32")
        # the module refuses to build its engine off-GPU; hand it one on the stand-ins (test only)
        model._engine = eng = engine.MegaEngine(model.state_dict(), engine_config_from(model.cfg), "cpu")
        eng.use_graph = False
        common = {"seg_len": total, "pattern": "%06d", "img_dir": "/nonexistent/%s.JPEG"}
        out = model({"cur": frames[0][0], "ref_l": [], "ref_g": [frames[j][0] for j in gpf[0]], "frame_category": 0,
                     "lookahead": [f[0] for f in frames[1:13]], **common})
        ref = gold["frames"][0]
        assert len(out) == 1 and out[0].size == (w, h)
        assert torch.equal(out[0].get_field("labels"), ref["labels"])
        assert torch.allclose(out[0].bbox, ref["boxes"], atol=2e-2) and torch.allclose(out[0].get_field("scores"), ref["scores"], atol=1e-3)
        _check_frame(eng, eng_det(eng), ref, 0)
        out = model({"cur": frames[1][0], "ref_l": [frames[min(1 + 12, total - 1)][0]], "ref_g": [frames[gpf[1][0]][0]],
                     "frame_category": 1, **common})
        ref = gold["frames"][1]
        assert torch.equal(out[0].get_field("labels"), ref["labels"]) and torch.allclose(out[0].bbox, ref["boxes"], atol=2e-2)
        _check_frame(eng, eng_det(eng), ref, 1)


def eng_det(eng):
    """the engine's static detection buffers of the last frame"""
    from mega_core.b200.engine import Detections
    b = eng._bufs
    get = lambda tag: [v for (t, _, _), v in b.items() if t == tag][0]        # noqa: E731
    return Detections(get("det_boxes"), get("det_scores"), get("det_labels"), get("det_count"))


def test_base_engine_logic_matches_reference_fixture():
    """single-frame R-50-C4 (BASELINE configs[0]) incl. the channel-reduction conv"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "base_r50_192x320.pt"))
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    with cpu_ops():
        eng = engine.BaseEngine(sd, engine.EngineConfig(precision="tf32"), device="cpu")
        det = eng.forward(synth.synthetic_frame(gold["frame_index"], gold["h"], gold["w"]), gold["w"], gold["h"])
        _check_frame(eng, det, gold, 0)


def test_rdn_engine_logic_matches_reference_fixture():
    """RdnEngine: 37-frame window filled by start_video, base stages + advanced stage"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "rdn_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    with cpu_ops():
        eng = engine.RdnEngine(sd, engine.EngineConfig(all_frame_interval=37, key_frame_location=18, stage=2,
                                                       advanced_stage=1, precision="tf32"), device="cpu")
        eng.use_graph = False
        det = eng.start_video(frames[0], frames[1:19], w, h)
        _check_frame(eng, det, gold["frames"][0], 0)


def test_fgfa_engine_logic_matches_reference_fixture():
    """FgfaEngine: frame / image rings, 19 FlowNetS pairs, warp + adaptive weights + aggregation"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "fgfa_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    with cpu_ops():
        eng = engine.FgfaEngine(sd, engine.EngineConfig(all_frame_interval=19, key_frame_location=9, precision="tf32"),
                                device="cpu")
        det = eng.start_video(frames[0], frames[1:10], w, h)
        ref = gold["frames"][0]
        assert (eng.last_flow[..., :2].permute(0, 3, 1, 2) - ref["flow"]).abs().max() < 1e-4
        _check_frame(eng, det, ref, 0)


def test_base_and_dff_module_api_on_standins():
    """GeneralizedRCNN (list-of-images call) and GeneralizedRCNNDFF ({cur, is_key_frame} dict) through
    build_detection_model / load_state_dict / model(...) -> list[BoxList], against the reference fixtures"""
    from mega_core.b200 import engine, synth
    from mega_core.modeling.detector import build_detection_model_from_state_dict
    from mega_core.modeling.nets import engine_config_from
    with cpu_ops():
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "base_r50_192x320.pt"))
        sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
        model = build_detection_model_from_state_dict(sd, method="base", device="cpu", precision="tf32")
        model._engine = engine.BaseEngine(model.state_dict(), engine_config_from(model.cfg), "cpu")
        out = model([synth.synthetic_frame(gold["frame_index"], gold["h"], gold["w"])[0]])
        assert torch.equal(out[0].get_field("labels"), gold["labels"]) and torch.allclose(out[0].bbox, gold["boxes"], atol=2e-2)
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "dff_r101_192x320.pt"))
        sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
        model = build_detection_model_from_state_dict(sd, method="dff", device="cpu", precision="tf32")
        model._engine = engine.DffEngine(model.state_dict(), engine_config_from(model.cfg), "cpu")
        for t in range(2):
            out = model({"cur": synth.synthetic_frame(gold["frame_stride"] * t, gold["h"], gold["w"])[0],
                         "is_key_frame": gold["key_flags"][t]})
            ref = gold["frames"][t]
            assert torch.equal(out[0].get_field("labels"), ref["labels"]) and torch.allclose(out[0].bbox, ref["boxes"], atol=2e-2)


def test_two_key_frames_per_call_equal_two_calls():
    """MegaEngine.step2_batched (the per-frame branch of two key frames as one batch of four images, then the two
    aggregations in order) gives exactly what two step_batched calls give -- detections, predictor rows, every ring"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "mega_r101_192x320.pt"))
    h, w = 96, 160
    sd = synth.make_state_dict("mega_r101_tiny", seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(40)]
    snap = lambda e, d: (e.last_pred[:int(e.cur_cnt.view(-1)[0])].clone(), d.boxes[:int(d.count[0])].clone(),   # noqa: E731
                         d.labels[:int(d.count[0])].clone())
    with cpu_ops():
        a = engine.MegaEngine(sd, engine.EngineConfig(precision="tf32"), device="cpu")
        b = engine.MegaEngine(sd, engine.EngineConfig(precision="tf32"), device="cpu")
        from mega_core.b200 import parallel
        for e in (a, b):
            e.use_graph = False
            parallel.random_state(e, 4, w, h)          # a full window / global pool without 23 backbone passes
        for t in range(1, 5, 2):
            quad = torch.cat([frames[t + 12], frames[30 + t], frames[t + 13], frames[31 + t]], 0)
            one = [snap(a, a.step_batched(quad[0:2], w, h)), snap(a, a.step_batched(quad[2:4], w, h))]
            d0, d1 = b.step2_batched(quad, w, h)
            k0 = one[0][0].shape[0]
            assert torch.equal(d0.boxes[:int(d0.count[0])], one[0][1]) and torch.equal(d0.labels[:int(d0.count[0])], one[0][2])
            two1 = snap(b, d1)
            for x, y in zip(one[1], two1):
                assert torch.equal(x, y)
            assert k0 > 0 and one[1][1].shape[0] > 0
        for name in ("E0", "B0", "Y1E", "Y2M", "B1", "B2", "win_x", "win_boxes", "win_cnt", "glob_x"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
