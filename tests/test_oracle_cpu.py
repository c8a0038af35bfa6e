"""CPU tests: the oracle (oracle/mega_oracle.py + oracle/csrc/oracle_ops.c) against the fixtures that
oracle/make_golden.py produced from the UNMODIFIED reference (reference unit-test vectors, its
compiled CPU ops, and end-to-end runs of its Python model)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")

import mega_oracle as mo  # noqa: E402


def _synth():
    from mega_core.b200 import synth
    return synth


def _same_boxes(a, b, atol=2e-3):
    """same box lists (the keep decisions are integer work and must agree), coordinates to 2e-3 px: the end-to-end
    fixtures were written on another host, and the oracle's convolutions are PyTorch's CPU kernels whose summation
    order depends on the CPU's ISA (third-party arithmetic, unpinned); on the machine that
    generated them oracle/make_golden.py shows max |diff| = 0 against the reference."""
    return a.shape == b.shape and torch.allclose(a, b, atol=atol, rtol=0)


def test_nms_reference_golden_vectors():
    """tests/test_nms.py:16-58 and :65-217 of the reference (6 cases)"""
    gold = torch.load(os.path.join(GOLD, "reference_unit_vectors.pt"))
    assert len(gold["nms"]) == 6
    for case in gold["nms"]:
        for sem in (False, True):
            keep = mo.nms(case["boxes"], case["scores"], case["thresh"], cuda_semantics=sem)
            assert keep.tolist() == sorted(case["expected"].tolist())
            assert torch.equal(keep, case["keep"].sort()[0])


def test_box_decode_reference_golden_vector():
    """tests/test_box_coder.py:15-105 of the reference (atol 1e-4 there; bit-exact vs its decode here)"""
    d = torch.load(os.path.join(GOLD, "reference_unit_vectors.pt"))["decode"]
    out = mo.decode_boxes(d["deltas"], d["boxes"], d["weights"])
    assert torch.allclose(out, d["expected"].float(), atol=1e-4)
    assert torch.equal(out, d["out"])


def test_anchors_match_reference_and_comment_table():
    gold = torch.load(os.path.join(GOLD, "reference_ops.pt"))
    assert torch.equal(mo.cell_anchors(16, (64, 128, 256, 512), (0.5, 1.0, 2.0)), gold["cell_anchors"])
    # The 9-anchor table in the reference's comments (rpn/anchor_generator.py:199-217) is the original
    # MATLAB/py-faster-rcnn one (1-based pixel coordinates): the reference's own generate_anchors returns
    # exactly that table minus 1 for sizes (128, 256, 512).
    table = torch.tensor([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                          [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                          [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], dtype=torch.float32)
    assert torch.equal(mo.cell_anchors(16, (128, 256, 512), (0.5, 1, 2)) + 1, table)
    g = mo.grid_anchors(3, 5, 16)
    assert g.shape == (3 * 5 * 12, 4)
    assert torch.equal(g[12], g[0] + torch.tensor([16.0, 0.0, 16.0, 0.0]))


def test_roi_align_matches_compiled_reference():
    gold = torch.load(os.path.join(GOLD, "reference_ops.pt"))
    for case in gold["roi_align"]:
        c, h, w, k, sr = case["seed_case"]
        assert torch.equal(mo.roi_align(case["feat"], case["rois"], 1.0 / 16, 7, 7, sr), case["out"])
    assert mo.roi_align(torch.zeros(1, 4, 5, 5), torch.zeros(0, 5), 1 / 16, 7, 7, 0).shape == (0, 4, 7, 7)


def test_nms_random_matches_compiled_reference():
    gold = torch.load(os.path.join(GOLD, "reference_ops.pt"))
    for case in gold["nms_random"]:
        if case["boxes"] is None:
            continue
        assert torch.equal(mo.nms(case["boxes"], case["scores"], case["thr"], False), case["keep_cpu"])
        assert torch.equal(mo.nms(case["boxes"], case["scores"], case["thr"], True), case["keep_cuda_sem"])
    assert mo.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).numel() == 0


def test_nms_tie_semantics_differ_only_on_exact_threshold():
    """nms_cpu.cpp:60 (>=) vs nms.cu:60 (>): IoU exactly 0.5 is suppressed only by the CPU rule"""
    boxes = torch.tensor([[0.0, 0.0, 9.0, 9.0], [0.0, 0.0, 9.0, 4.0]])   # areas 100 / 50, inter 50 -> IoU 0.5
    scores = torch.tensor([0.9, 0.8])
    assert mo.nms(boxes, scores, 0.5, cuda_semantics=False).tolist() == [0]
    assert mo.nms(boxes, scores, 0.5, cuda_semantics=True).tolist() == [0, 1]


def test_position_embedding_layout():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(3, 2, generator=g) * 100
    bq = torch.cat([a, a + 20], 1)
    pe = mo.position_embedding(bq, bq)
    assert pe.shape == (64, 3, 3)
    # identical boxes: dw = dh = 0 -> sin 0 / cos 1 in channels 32..63
    assert torch.allclose(pe[32:40, 0, 0], torch.zeros(8)) and torch.allclose(pe[40:48, 0, 0], torch.ones(8))


def test_base_r50_oracle_matches_reference_fixture():
    synth = _synth()
    gold = torch.load(os.path.join(GOLD, "base_r50_192x320.pt"))
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    orc = mo.BaseOracle(sd, record=True)
    b, s, l = orc.forward(synth.synthetic_frame(gold["frame_index"], gold["h"], gold["w"]))
    assert torch.allclose(orc.trace["class_logits"], gold["class_logits"], atol=1e-5)
    assert torch.equal(l, gold["labels"]) and _same_boxes(b, gold["boxes"]) and torch.allclose(s, gold["scores"], atol=1e-6)


def test_mega_r101_oracle_matches_reference_fixture():
    """4 frames of the unmodified reference's GeneralizedRCNNMEGA (memory filling from empty)"""
    synth = _synth()
    gold = torch.load(os.path.join(GOLD, "mega_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    orc = mo.MegaOracle(sd, record=True)
    gpf = gold["globals_per_frame"]
    for t, ref in enumerate(gold["frames"][:3]):
        infos = {"frame_category": 0 if t == 0 else 1,
                 "ref_l": frames[1:13] if t == 0 else [frames[min(t + 12, total - 1)]],
                 "ref_g": [frames[j] for j in gpf[t]]}
        b, s, l = orc.forward(frames[t], infos)
        assert torch.allclose(orc.trace["class_logits"], ref["class_logits"], atol=1e-5)
        assert _same_boxes(orc.trace["proposals"], ref["proposals"])
        assert torch.equal(l, ref["labels"]) and _same_boxes(b, ref["boxes"])


def test_rdn_r101_oracle_matches_reference_fixture():
    """2 frames of the unmodified reference's GeneralizedRCNNRDN (37-frame window, base stages + advanced stage)"""
    synth = _synth()
    gold = torch.load(os.path.join(GOLD, "rdn_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(total)]
    orc = mo.RdnOracle(sd, record=True)
    for t, ref in enumerate(gold["frames"][:2]):
        infos = {"frame_category": 0 if t == 0 else 1, "ref": frames[1:19] if t == 0 else [frames[min(t + 18, total - 1)]]}
        b, s, l = orc.forward(frames[t], infos)
        assert torch.allclose(orc.trace["class_logits"], ref["class_logits"], atol=1e-5)
        assert _same_boxes(orc.trace["proposals"], ref["proposals"])
        assert torch.equal(l, ref["labels"]) and _same_boxes(b, ref["boxes"])


def test_fgfa_r101_oracle_matches_reference_fixture():
    """frame 0 of the unmodified reference's GeneralizedRCNNFGFA (FlowNetS + EmbedNet + warp / weights / aggregation)"""
    synth = _synth()
    gold = torch.load(os.path.join(GOLD, "fgfa_r101_192x320.pt"))
    h, w, total = gold["h"], gold["w"], gold["total"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    frames = [synth.synthetic_frame(i, h, w) for i in range(10)]
    orc = mo.FgfaOracle(sd, record=True)
    ref = gold["frames"][0]
    b, s, l = orc.forward(frames[0], {"frame_category": 0, "ref": frames[1:10]})
    assert torch.allclose(orc.trace["flow"], ref["flow"], atol=1e-5)
    assert torch.allclose(orc.trace["class_logits"], ref["class_logits"], atol=1e-5)
    assert _same_boxes(orc.trace["proposals"], ref["proposals"]) and torch.equal(l, ref["labels"])


def test_dff_r101_oracle_matches_reference_fixture():
    """5 frames (key, 2 non-key, key, non-key) of the unmodified reference's GeneralizedRCNNDFF: FlowNetS flow + scale
    map, warped key-frame features, single-frame head"""
    synth = _synth()
    gold = torch.load(os.path.join(GOLD, "dff_r101_192x320.pt"))
    h, w = gold["h"], gold["w"]
    sd = synth.make_state_dict(gold["arch"], seed=gold["seed"])
    orc = mo.DffOracle(sd, record=True)
    for t, (key, ref) in enumerate(zip(gold["key_flags"], gold["frames"])):
        if t >= 3:
            break                                                      # keep the CPU suite short
        b, s, l = orc.forward(synth.synthetic_frame(gold["frame_stride"] * t, h, w), key)
        assert torch.allclose(orc.trace["flow"], ref["flow"], atol=1e-5)
        assert torch.allclose(orc.trace["scale"][:, ::64], ref["scale_sample"], atol=1e-5)
        assert torch.allclose(orc.trace["class_logits"], ref["class_logits"], atol=1e-5)
        assert _same_boxes(orc.trace["proposals"], ref["proposals"]) and torch.equal(l, ref["labels"])
