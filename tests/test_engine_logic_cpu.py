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
