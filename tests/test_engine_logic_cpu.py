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


def test_mega_engine_logic_matches_reference_fixture():
    """MegaEngine (the hot path: start_video with 13 local + 10 global frames, then a steady-state step through the
    index tables, window / global rings, long-range memory pushes, fused aggregation) vs the unmodified reference's
    GeneralizedRCNNMEGA outputs"""
    from mega_core.b200 import engine, synth
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "mega_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    gpf = gold["globals_per_frame"]
    with cpu_ops():
        eng = engine.MegaEngine(sd, engine.EngineConfig(precision="tf32"), device="cpu")
        eng.use_graph = False
        det = eng.start_video(frames[0], frames[1:13], [frames[j] for j in gpf[0]], w, h)
        _check_frame(eng, det, gold["frames"][0], 0)
        det = eng.step(frames[min(1 + 12, total - 1)], frames[gpf[1][0]], w, h)
        _check_frame(eng, det, gold["frames"][1], 1)


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
