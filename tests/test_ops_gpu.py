"""GPU parity of the non-GEMM kernels against the oracle (oracle/mega_oracle.py, oracle_ops.c).

Index outputs (NMS keeps, proposal selection given identical scores) are compared bit-exactly;
fp32 arithmetic that uses only +,-,*,/ in the reference's association order (IoU, ROIAlign) is
compared bit-exactly too; kernels that call exp/log/sin/cos are compared within a few ulp-scale
tolerances stated at each assert.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _oracle():
    import mega_oracle
    return mega_oracle


def _rand_boxes(n, g, w=1000.0, h=600.0, lo=4.0, hi=260.0):
    xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.9, h * 0.9])
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    b = torch.cat([xy, xy + wh], 1)
    b[:, 0::2].clamp_(0, w - 1)
    b[:, 1::2].clamp_(0, h - 1)
    return b


@pytest.mark.parametrize("n,thr", [(1, 0.5), (5, 0.3), (64, 0.5), (65, 0.7), (300, 0.5), (2000, 0.7), (6000, 0.7),
                                   (8192, 0.6)])
def test_nms_matches_oracle_bitexact(cuda_dev, n, thr):
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(n)
    boxes = _rand_boxes(n, g)
    scores = torch.rand(n, generator=g)
    if n >= 64:  # exact score ties and duplicate boxes exercise the tie rule / the strict '>'
        scores[10] = scores[3]
        boxes[20] = boxes[7]
    keep, cnt = ops.nms_device(boxes.to(cuda_dev), scores.to(cuda_dev), thr)
    got = keep[: int(cnt.item())].cpu()
    ref = mo.nms(boxes, scores, thr, cuda_semantics=True)
    assert torch.equal(got, ref)


def test_nms_reference_golden_vectors(cuda_dev):
    """the reference's own known-answer vectors (tests/test_nms.py:16-58, :65-217)"""
    from mega_core.b200 import ops
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_unit_vectors.pt"))
    for case in gold["nms"]:
        keep, cnt = ops.nms_device(case["boxes"].to(cuda_dev), case["scores"].to(cuda_dev), case["thresh"])
        assert keep[: int(cnt.item())].cpu().tolist() == sorted(case["expected"].tolist())


def test_nms_empty(cuda_dev):
    from mega_core.b200 import ops
    keep, cnt = ops.nms_device(torch.zeros(0, 4, device=cuda_dev), torch.zeros(0, device=cuda_dev), 0.5)
    assert int(cnt.item()) == 0


@pytest.mark.parametrize("c,h,w,k,sr", [(8, 20, 30, 24, 0), (16, 13, 17, 9, 2), (2048, 38, 63, 75, 0), (256, 38, 63, 300, 0)])
def test_roi_align_bitexact(cuda_dev, c, h, w, k, sr):
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(c + k)
    feat = torch.randn(1, c, h, w, generator=g)
    boxes = _rand_boxes(k, g, w * 16.0, h * 16.0, 1.0, w * 10.0)
    boxes[0] = torch.tensor([5.0, 5.0, 5.0, 5.0])
    boxes[1] = torch.tensor([-50.0, -60.0, 3000.0, 2000.0])
    rois = torch.cat([torch.zeros(k, 1), boxes], 1)
    ref = mo.roi_align(feat, rois, 1.0 / 16, 7, 7, sr)                      # [K,C,7,7]
    got_nchw = ops.roi_align_nchw(feat.to(cuda_dev), rois.to(cuda_dev), 1.0 / 16, 7, 7, sr).cpu()
    assert torch.equal(got_nchw, ref), (got_nchw - ref).abs().max()
    nhwc = feat.permute(0, 2, 3, 1).contiguous().to(cuda_dev)
    out = torch.full((k, 49 * c), float("nan"), device=cuda_dev)
    ops.roi_align_nhwc(nhwc, boxes.to(cuda_dev), None, 1.0 / 16, 7, 7, sr, out)
    got = out.view(k, 49, c).permute(0, 2, 1).reshape(k, c, 7, 7).cpu()
    assert torch.equal(got, ref), (got - ref).abs().max()


def test_roi_align_golden_fixture(cuda_dev):
    """outputs of the reference's compiled ROIAlign_cpu.cpp committed by oracle/make_golden.py"""
    from mega_core.b200 import ops
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_ops.pt"))
    for case in gold["roi_align"]:
        c, h, w, k, sr = case["seed_case"]
        got = ops.roi_align_nchw(case["feat"].to(cuda_dev), case["rois"].to(cuda_dev), 1.0 / 16, 7, 7, sr).cpu()
        assert torch.equal(got, case["out"])


def test_stem_maxpool_gather_transpose(cuda_dev):
    import torch.nn.functional as F
    from mega_core.b200 import ops
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 37, 53, generator=g)
    ho, wo = 19, 27
    col = torch.empty(2, ho * wo, 160, device=cuda_dev)
    ops.stem_im2col(img.to(cuda_dev), col)
    ref = F.unfold(img, 7, padding=3, stride=2).transpose(1, 2)              # [2, ho*wo, 147] (c, r, s) order
    assert torch.equal(col[:, :, :147].cpu(), ref) and (col[:, :, 147:] == 0).all()
    x = torch.randn(2, 64, 21, 30, generator=g)
    out = torch.empty(2, 11, 15, 64, device=cuda_dev)
    ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous().to(cuda_dev), out)
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), F.max_pool2d(x, 3, 2, 1))
    src = torch.randn(50, 1024, generator=g)
    idx = torch.tensor([3, 3, -1, 49, 0, 17], dtype=torch.int32)
    dst = torch.full((6, 1024), 7.0, device=cuda_dev)
    ops.gather_rows(src.to(cuda_dev), idx.to(cuda_dev), dst)
    exp = src[idx.clamp_min(0).long()]
    exp[2] = 0
    assert torch.equal(dst.cpu(), exp)
    dst2 = torch.zeros(60, 1024, device=cuda_dev)
    didx = torch.tensor([59, 0, 7], dtype=torch.int32, device=cuda_dev)
    ops.copy_rows(src[:3].to(cuda_dev), dst2, 3, dst_idx=didx)
    assert torch.equal(dst2[[59, 0, 7]].cpu(), src[:3])
    cnt = torch.tensor([123, 456], dtype=torch.int32, device=cuda_dev)
    ring = torch.zeros(25, 1, dtype=torch.int32, device=cuda_dev)
    ops.copy_rows(cnt[1:2].view(torch.float32).view(1, 1), ring.view(torch.float32), 1, row_len=1,
                  dst_idx=torch.tensor([9], dtype=torch.int32, device=cuda_dev))
    assert ring[9, 0].item() == 456 and ring.sum().item() == 456
    m = torch.randn(3, 40, 70, generator=g)
    t = torch.empty(3, 70, 40, device=cuda_dev)
    ops.transpose_2d(m.to(cuda_dev), t, 3, 40, 70)
    assert torch.equal(t.cpu(), m.transpose(1, 2))


@pytest.mark.parametrize("h,w,post", [(12, 20, 300), (38, 63, 300), (38, 63, 75)])
def test_rpn_select_matches_oracle(cuda_dev, h, w, post):
    """identical head outputs in -> proposals out. Scores go through sigmoid (exp): the GPU top-k is
    checked to be a valid ordering of the oracle's scores within 2 ulp, and -- given the GPU's own
    pre-NMS boxes -- the NMS keep list must equal the oracle's bit-exactly."""
    from mega_core.b200 import engine, ops
    mo = _oracle()
    g = torch.Generator().manual_seed(h * w)
    a = 12
    logits = torch.randn(1, a, h, w, generator=g) * 1.5
    deltas = torch.randn(1, 4 * a, h, w, generator=g) * 0.3
    im_w, im_h = w * 16.0 - 8, h * 16.0 - 8
    head = torch.zeros(1, h, w, 64)
    head[0, :, :, :a] = logits[0].permute(1, 2, 0)
    head[0, :, :, a:5 * a] = deltas[0].permute(1, 2, 0)
    base = engine.cell_anchors(16, (64, 128, 256, 512), (0.5, 1.0, 2.0))
    boxes, scores, anchor, cnt = ops.rpn_select(head.to(cuda_dev), 1, h, w, base.to(cuda_dev), im_w, im_h, 6000, post,
                                                0.7, 0.0, 16, want_anchor=True)
    n = int(cnt[0].item())
    rb, rs, aux = mo.rpn_select(logits, deltas, im_w, im_h, 6000, post, 0.7, cuda_semantics=True, return_aux=True)
    assert n == rb.shape[0]
    got_anchor = anchor[0, :n].cpu().long()
    assert torch.equal(got_anchor, aux["anchor_idx"]), "selected anchors differ from the oracle"
    assert torch.allclose(boxes[0, :n].cpu(), rb, rtol=0, atol=2e-3)        # exp() in the decode: few ulp at 1e3 px
    assert torch.allclose(scores[0, :n].cpu(), rs, rtol=0, atol=2e-7)
    assert (boxes[0, n:] == 0).all()


def test_relation_softmax_matches_oracle(cuda_dev):
    """position-embedding bias + soft-max vs the oracle's materialised [64,N,M] path (fp32 tolerance:
    1e-5 absolute on probabilities that sum to one; sin/cos arguments reach ~700 rad)."""
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(5)
    n, m, ld = 37, 203, 224
    bq, bk = _rand_boxes(n, g), _rand_boxes(m, g)
    wg = torch.randn(16, 64, generator=g) * 0.2
    bg = torch.rand(16, generator=g) * 0.5
    aff = torch.randn(16, n, m, generator=g) * 8.0
    pe = mo.position_embedding(bq, bk)                                       # [64, n, m]
    w = torch.relu(torch.einsum("ge,enm->gnm", wg, pe) + bg.view(16, 1, 1))
    ref = torch.softmax((w + 1e-6).log() + aff * 0.125, dim=2)
    s = torch.zeros(16, n, ld, device=cuda_dev)
    s[:, :, :m] = aff.to(cuda_dev)
    s[:, :, m:] = 123.0
    dim_mat = torch.full((8,), 1000.0).pow(8.0 / 64 * torch.arange(0, 8, dtype=torch.float32))
    mv = torch.tensor([m], dtype=torch.int32, device=cuda_dev)
    ops.relation_softmax(s, n, ld, 0.125, boxes_q=bq.to(cuda_dev), boxes_k=bk.to(cuda_dev), wg=wg.to(cuda_dev),
                         bg=bg.to(cuda_dev), dim_mat=dim_mat.to(cuda_dev), m_valid=mv)
    got = s.cpu()
    assert (got[:, :, m:] == 0).all()
    assert (got[:, :, :m] - ref).abs().max() < 1e-5
    # without the position term, with padded query rows skipped
    s2 = torch.zeros(16, n, ld, device=cuda_dev)
    s2[:, :, :m] = aff.to(cuda_dev)
    nv = torch.tensor([30], dtype=torch.int32, device=cuda_dev)
    ops.relation_softmax(s2, n, ld, 0.125, m_host=m, n_valid=nv, n_valid_off=35)
    ref2 = torch.softmax(aff * 0.125, dim=2)
    got2 = s2.cpu()
    live = [i for i in range(n) if i < 30 or i >= 35]
    assert (got2[:, live, :m] - ref2[:, live]).abs().max() < 1e-6
    assert torch.equal(got2[:, 30:35, :m], aff[:, 30:35])


def test_box_postprocess_matches_oracle(cuda_dev):
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(9)
    for r, k_valid, scale in ((300, 300, 0.3), (300, 137, 2.5), (40, 40, 4.0)):
        logits = torch.randn(r, 31, generator=g) * scale
        deltas = torch.randn(r, 124, generator=g) * 0.5
        props = _rand_boxes(r, g)
        pred = torch.zeros(r, 156)
        pred[:, :31] = logits
        pred[:, 31:155] = deltas
        pd = pred.to(cuda_dev)
        cap = 30 * r
        out = (torch.zeros(cap, 4, device=cuda_dev), torch.zeros(cap, device=cuda_dev),
               torch.zeros(cap, dtype=torch.int64, device=cuda_dev), torch.zeros(1, dtype=torch.int32, device=cuda_dev))
        cnt = torch.tensor([k_valid], dtype=torch.int32, device=cuda_dev)
        ops.box_postprocess(pd[:, :31], pd[:, 31:], props.to(cuda_dev), cnt, 31, 1000.0, 600.0, 0.001, 0.5, 300,
                            (10.0, 10.0, 5.0, 5.0), out)
        n = int(out[3].item())
        rb, rs, rl = mo.box_postprocess(logits[:k_valid], deltas[:k_valid], props[:k_valid], 1000.0, 600.0,
                                        cuda_semantics=True)
        assert n == rb.shape[0], (n, rb.shape)
        assert torch.equal(out[2][:n].cpu(), rl)
        assert torch.allclose(out[1][:n].cpu(), rs, rtol=0, atol=1e-6)    # 31-way softmax: exp + sum order
        assert torch.allclose(out[0][:n].cpu(), rb, rtol=0, atol=2e-3)


def test_f16_variants_of_memory_bound_kernels(cuda_dev):
    """fp16-storage variants used by the fp16-operand engine: same fp32 arithmetic as the fp32 kernels on
    fp16-representable inputs, one rounding to fp16 at the store (exact for im2col / max-pool / row moves)."""
    import torch.nn.functional as F
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(11)
    # ROIAlign: fp16 map in -> fp16 out == fp16(round) of the oracle on the same (fp16-valued) map
    for c, h, w, k in ((2048, 38, 63, 75), (256, 20, 30, 40), (8, 13, 17, 9)):
        feat = torch.randn(1, c, h, w, generator=g).half()
        boxes = _rand_boxes(k, g, w * 16.0, h * 16.0, 1.0, w * 10.0)
        boxes[0] = torch.tensor([-50.0, -60.0, 3000.0, 2000.0])
        rois = torch.cat([torch.zeros(k, 1), boxes], 1)
        ref = mo.roi_align(feat.float(), rois, 1.0 / 16, 7, 7, 0).half()
        nhwc = feat.permute(0, 2, 3, 1).contiguous().to(cuda_dev)
        out = torch.full((k, 49 * c), float("nan"), device=cuda_dev, dtype=torch.float16)
        ops.roi_align_nhwc(nhwc, boxes.to(cuda_dev), None, 1.0 / 16, 7, 7, 0, out)
        got = out.view(k, 49, c).permute(0, 2, 1).reshape(k, c, 7, 7).cpu()
        if c % 128 == 0:     # fast path: fused multiply-adds, separable weights -> within one fp16 ulp of the oracle
            err = (got.float() - ref.float()).abs()
            assert (err <= ref.float().abs() * 2.0 ** -10 + 1e-6).all(), err.max()
        else:
            assert torch.equal(got, ref), (got.float() - ref.float()).abs().max()
    # stem im2col / max-pool
    img = torch.randn(2, 3, 37, 53, generator=g)
    col = torch.empty(2, 19 * 27, 160, device=cuda_dev, dtype=torch.float16)
    ops.stem_im2col(img.to(cuda_dev), col)
    ref = F.unfold(img, 7, padding=3, stride=2).transpose(1, 2).half()
    assert torch.equal(col[:, :, :147].cpu(), ref) and (col[:, :, 147:] == 0).all()
    x = torch.randn(2, 64, 21, 30, generator=g).half()
    out = torch.empty(2, 11, 15, 64, device=cuda_dev, dtype=torch.float16)
    ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous().to(cuda_dev), out)
    assert torch.equal(out.permute(0, 3, 1, 2).cpu().float(), F.max_pool2d(x.float(), 3, 2, 1))
    # row gathers over fp16 feature rows
    src = torch.randn(50, 1024, generator=g).half()
    idx = torch.tensor([3, 3, -1, 49, 0, 17], dtype=torch.int32)
    dst = torch.full((6, 1024), 7.0, device=cuda_dev, dtype=torch.float16)
    ops.gather_rows(src.to(cuda_dev), idx.to(cuda_dev), dst)
    exp = src[idx.clamp_min(0).long()]
    exp[2] = 0
    assert torch.equal(dst.cpu(), exp)
    # soft-max with fp16 probabilities: same values as the fp32 kernel, rounded once
    n, m, ld = 37, 203, 224
    bq, bk = _rand_boxes(n, g), _rand_boxes(m, g)
    wg = torch.randn(16, 64, generator=g) * 0.2
    bg = torch.rand(16, generator=g) * 0.5
    aff = torch.randn(16, n, m, generator=g) * 8.0
    dim_mat = torch.full((8,), 1000.0).pow(8.0 / 64 * torch.arange(0, 8, dtype=torch.float32))
    mv = torch.tensor([m], dtype=torch.int32, device=cuda_dev)
    outs = []
    for use16 in (False, True):
        s = torch.zeros(16, n, ld, device=cuda_dev)
        s[:, :, :m] = aff.to(cuda_dev)
        p16 = torch.full((16, n, ld), float("nan"), device=cuda_dev, dtype=torch.float16) if use16 else None
        ops.relation_softmax(s, n, ld, 0.125, boxes_q=bq.to(cuda_dev), boxes_k=bk.to(cuda_dev), wg=wg.to(cuda_dev),
                             bg=bg.to(cuda_dev), dim_mat=dim_mat.to(cuda_dev), m_valid=mv, probs_f16=p16)
        outs.append((p16 if use16 else s).cpu())
    assert torch.equal(outs[1], outs[0].half())
    s = torch.zeros(16, n, ld, device=cuda_dev)
    s[:, :, :m] = aff.to(cuda_dev)
    p16 = torch.full((16, n, ld), float("nan"), device=cuda_dev, dtype=torch.float16)
    ops.relation_softmax(s, n, ld, 0.125, m_host=m, probs_f16=p16)
    assert (p16.float().cpu()[:, :, :m] - torch.softmax(aff * 0.125, dim=2)).abs().max() < 5e-4
    assert (p16[:, :, m:] == 0).all()


@pytest.mark.parametrize("n,m,ld", [(37, 203, 224), (9, 1500, 1504)])
def test_relation_softmax_const_operand_kernel_matches_oracle(cuda_dev, n, m, ld):
    """the kernel-parameter (constant-bank) variant of the position-biased soft-max, both the shared-memory staged
    form (ld <= 1024) and the global two-pass form, fp32 in place and fp16 probabilities, vs the oracle"""
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(n + m)
    bq, bk = _rand_boxes(n, g), _rand_boxes(m, g)
    wg = torch.randn(16, 64, generator=g) * 0.2
    bg = torch.rand(16, generator=g) * 0.5
    aff = torch.randn(16, n, m, generator=g) * 8.0
    pe = mo.position_embedding(bq, bk)
    w = torch.relu(torch.einsum("ge,enm->gnm", wg, pe) + bg.view(16, 1, 1))
    ref = torch.softmax((w + 1e-6).log() + aff * 0.125, dim=2)
    dim_mat = torch.full((8,), 1000.0).pow(8.0 / 64 * torch.arange(0, 8, dtype=torch.float32)).contiguous()
    mv = torch.tensor([m], dtype=torch.int32, device=cuda_dev)
    for use16 in (False, True):
        s = torch.zeros(16, n, ld, device=cuda_dev)
        s[:, :, :m] = aff.to(cuda_dev)
        s[:, :, m:] = 123.0
        p16 = torch.full((16, n, ld), float("nan"), device=cuda_dev, dtype=torch.float16) if use16 else None
        ops.relation_softmax(s, n, ld, 0.125, boxes_q=bq.to(cuda_dev), boxes_k=bk.to(cuda_dev), m_valid=mv,
                             probs_f16=p16, host_w=(wg.contiguous(), bg.contiguous(), dim_mat))
        got = (p16.float() if use16 else s).cpu()
        assert (got[:, :, m:] == 0).all()
        assert (got[:, :, :m] - ref).abs().max() < (5e-4 if use16 else 1e-5)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_fgfa_kernels_match_oracle(cuda_dev, dtype):
    """csrc/fgfa.cu against the oracle's restatement of generalized_rcnn_fgfa.py:45-76, :198-214 and flownet.py:52-55"""
    import torch.nn.functional as F
    from mega_core.b200 import ops
    mo = _oracle()
    g = torch.Generator().manual_seed(8)
    L, key, h, w = 5, 2, 37, 54
    imgs = [torch.rand(1, 3, h, w, generator=g) * 255 - 110 for _ in range(L)]
    hq, wq = (h + 1) // 2, (w + 1) // 2
    ring = torch.zeros(L + 2, hq, wq, 4, device=cuda_dev, dtype=dtype)
    slots = [4, 6, 1, 0, 3]
    for im, s in zip(imgs, slots):
        ops.fgfa_pool_image(im.to(cuda_dev), ring[s])
    pooled_ref = [F.avg_pool2d(im / 255, 2, stride=2, ceil_mode=True) for im in imgs]
    tol = 1e-3 if dtype == torch.float16 else 1e-6
    for im, s in zip(pooled_ref, slots):
        assert (ring[s, :, :, :3].float().cpu().permute(2, 0, 1) - im[0]).abs().max() < tol
        assert (ring[s, :, :, 3] == 0).all()
    slots_d = torch.tensor(slots, dtype=torch.int32, device=cuda_dev)
    pairs = torch.full((L, hq + 6, wq + 8, 8), float("nan"), device=cuda_dev, dtype=dtype)
    ops.fgfa_build_pairs(ring, slots_d, key, pairs)
    pc = pairs.float().cpu()
    for i in range(L):
        assert (pc[i, 3:3 + hq, 3:3 + wq, 0:3].permute(2, 0, 1) - pooled_ref[key][0]).abs().max() < tol
        assert (pc[i, 3:3 + hq, 3:3 + wq, 4:7].permute(2, 0, 1) - pooled_ref[i][0]).abs().max() < tol
    assert (pc[:, :3] == 0).all() and (pc[:, :, :3] == 0).all() and (pc[..., 3] == 0).all() and (pc[..., 7] == 0).all()
    # avg-pool NHWC ceil_mode
    x = torch.randn(2, 11, 75, 125, generator=g)
    xin = torch.zeros(2, 75, 125, 16, device=cuda_dev, dtype=dtype)
    xin[..., :11] = x.permute(0, 2, 3, 1).to(cuda_dev).to(dtype)
    out = torch.zeros(2, 38, 63, 16, device=cuda_dev, dtype=dtype)
    ops.avgpool2_nhwc(xin, out)
    ref = F.avg_pool2d(xin[..., :11].float().cpu().permute(0, 3, 1, 2), 2, stride=2, ceil_mode=True)
    assert (out[..., :11].float().cpu().permute(0, 3, 1, 2) - ref).abs().max() < (2e-3 if dtype == torch.float16 else 1e-6)
    # warp + weights + aggregate
    fh, fw, cf, ce = 12, 20, 64, 128
    feats = [torch.randn(1, cf + ce, fh, fw, generator=g) for _ in range(L)]
    flow = torch.randn(L, 2, fh, fw, generator=g) * 1.5
    flow[0, :, 0, 0] = torch.tensor([-30.0, 25.0])            # far outside: border clamp
    fring = torch.zeros(L + 2, fh, fw, cf + ce, device=cuda_dev, dtype=dtype)
    for f, s in zip(feats, slots):
        fring[s] = f[0].permute(1, 2, 0).to(cuda_dev).to(dtype)
    allf = torch.cat([fring[s].float().cpu().permute(2, 0, 1)[None] for s in slots], 0)
    warped = mo.fgfa_warp(allf, flow)
    wf, emb = torch.split(warped, (cf, ce), dim=1)
    en = emb / (torch.norm(emb, dim=1, keepdim=True) + 1e-10)
    ec = emb[key:key + 1] / (torch.norm(emb[key:key + 1], dim=1, keepdim=True) + 1e-10)
    wts = torch.softmax(torch.sum(en * ec, dim=1, keepdim=True), dim=0)
    ref = torch.sum(wts * wf, dim=0)                                        # [cf, fh, fw]
    flow_d = torch.zeros(L, fh, fw, 4, device=cuda_dev)
    flow_d[..., :2] = flow.permute(0, 2, 3, 1).to(cuda_dev)
    out = torch.zeros(fh, fw, cf, device=cuda_dev, dtype=dtype)
    wout = torch.zeros(L, fh, fw, device=cuda_dev)
    ops.fgfa_aggregate(fring, slots_d, key, flow_d, out, cf, ce, weights_out=wout)
    assert (wout.cpu() - wts[:, 0]).abs().max() < 1e-5
    assert (out.float().cpu().permute(2, 0, 1) - ref).abs().max() < (3e-3 if dtype == torch.float16 else 1e-5)


def test_copy_batch_matches_individual_copies(cuda_dev):
    """ops.copy_batch(): a group of independent gather / scatter row copies issued as one launch"""
    from mega_core.b200 import ops
    g = torch.Generator().manual_seed(2)
    src = torch.randn(60, 1024, generator=g).to(cuda_dev)
    src16 = torch.randn(40, 1024, generator=g).half().to(cuda_dev)
    boxes = torch.randn(60, 4, generator=g).to(cuda_dev)
    cnt = torch.tensor([[7], [9], [11]], dtype=torch.int32, device=cuda_dev)
    idx = torch.tensor([5, -1, 59, 0, 0, 33], dtype=torch.int32, device=cuda_dev)
    didx = torch.tensor([3, 1, -1, 0], dtype=torch.int32, device=cuda_dev)
    sel = torch.tensor([2], dtype=torch.int32, device=cuda_dev)

    def run(batched):
        outs = [torch.full((6, 1024), 5.0, device=cuda_dev), torch.full((6, 4), 5.0, device=cuda_dev),
                torch.full((8, 1024), 5.0, device=cuda_dev, dtype=torch.float16), torch.zeros(1, 1, dtype=torch.int32, device=cuda_dev)]
        ctx = ops.copy_batch() if batched else None
        if ctx:
            ctx.__enter__()
        ops.gather_rows(src, idx, outs[0])
        ops.gather_rows(boxes, idx, outs[1])
        ops.copy_rows(src16[:4], outs[2], 4, dst_idx=didx)
        ops.gather_rows(cnt.view(torch.float32), sel, outs[3].view(torch.float32), 1, row_len=1)
        if ctx:
            ctx.__exit__(None, None, None)
        torch.cuda.synchronize()
        return outs

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert int(b[3].item()) == 11
