"""CPU stand-ins for the device ops of mega_core.b200.ops -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the engines' HOST LOGIC (ring buffers, per-frame index tables, launch order, multi-GPU schedules) on
CPU tensors in the `-m "not gpu"` suite. Inside `with cpu_ops():` every `ops.*` entry point an engine calls is
replaced by a deterministic torch function with the same contract (same arguments, same in-place effects, same
masking rules), the dense contractions by the exact-fp32 shadow of tests/fp32_shadow.py, and the few CUDA runtime
calls the engines make on the host (pinned memory, events, streams) by inert stubs. Results are comparable BETWEEN runs
on the stand-ins (e.g. wavefront schedule vs sequential schedule) and, where a stand-in is built from the oracle's own
function, against the oracle; they say nothing about the CUDA kernels, which have their own parity tests.
Nothing in the product imports this module, and the product still has no CPU path: outside the context manager the
ops raise on CPU tensors as before.
"""
import contextlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import mega_oracle as mo  # noqa: E402
from fp32_shadow import _shadow_conv_gemm  # noqa: E402


def gather_rows(src, idx, dst, n_rows=None, row_len=None):
    n_rows = idx.numel() if n_rows is None else n_rows
    row_len = src.shape[-1] if row_len is None else row_len
    i = idx[:n_rows].long()
    val = src[i.clamp_min(0), :row_len].clone()
    val[i < 0] = 0
    dst[:n_rows, :row_len] = val
    return dst


def copy_rows(src, dst, n_rows, row_len=None, src_idx=None, dst_idx=None):
    row_len = src.shape[-1] if row_len is None else row_len
    si = src_idx[:n_rows].long() if src_idx is not None else torch.arange(n_rows)
    di = dst_idx[:n_rows].long() if dst_idx is not None else torch.arange(n_rows)
    val = src[si.clamp_min(0), :row_len].clone()
    val[si < 0] = 0
    keep = di >= 0
    dst[di[keep], :row_len] = val[keep]
    return dst


class copy_batch(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def relation_softmax(logits, n_rows, ldm, scale, boxes_q=None, boxes_k=None, wg=None, bg=None, dim_mat=None,
                     m_valid=None, m_host=0, n_valid=None, n_valid_off=0, probs_f16=None, host_w=None):
    """include/mega_b200.h mega_relation_softmax: p = softmax_m(log(relu(Wg.emb + bg) + 1e-6) + scale * logits) over the
    keys m < m_valid, probability 0 beyond; query rows n_valid <= n < n_valid_off are padding and stay untouched"""
    m = int(m_valid.reshape(-1)[0]) if m_valid is not None else int(m_host)
    s = logits[:, :n_rows, :m].double() * scale
    if boxes_q is not None:
        pe = mo.position_embedding(boxes_q[:n_rows].float(), boxes_k[:m].float()).double()          # [64, N, M]
        gate = F.relu(torch.einsum("gc,cnm->gnm", wg.double(), pe) + bg.double().view(-1, 1, 1))
        s = s + torch.log(gate + 1e-6)
    p = torch.softmax(s, dim=2).float()
    live = torch.ones(n_rows, dtype=torch.bool)
    if n_valid is not None:
        nv = int(n_valid.reshape(-1)[0])
        live[nv:n_valid_off] = False
    out = probs_f16 if probs_f16 is not None else logits
    full = torch.zeros(logits.shape[0], n_rows, logits.shape[2], dtype=out.dtype)
    full[:, :, :m] = p.to(out.dtype)
    out[:, :n_rows][:, live] = full[:, live]
    return out


def box_postprocess(logits, deltas, proposals, count, num_classes, im_w, im_h, score_thresh, nms_thresh, max_det,
                    weights, out):
    k = int(count.reshape(-1)[0])
    b, s, l = mo.box_postprocess(logits[:k, :num_classes].float(), deltas[:k, :4 * num_classes].float(), proposals[:k],
                                 im_w, im_h, score_thresh, nms_thresh, max_det, weights, cuda_semantics=True)
    ob, os_, ol, oc = out
    n = b.shape[0]
    ob.zero_(), os_.zero_(), ol.zero_()
    ob[:n], os_[:n], ol[:n] = b, s, l
    oc.fill_(n)
    return out


# ------------------------------------------------------------------- per-frame branch (backbone, proposals, ROIAlign)
def stem_prep(img, out):
    """NCHW fp32 image -> zero-bordered NHWC8 [N, H+6, wp, 8] (3 real channels)"""
    n, _, h, w = img.shape
    out.zero_()
    out[:, 3:3 + h, 3:3 + w, 0:3] = img.permute(0, 2, 3, 1).to(out.dtype)
    return out


def maxpool3x3s2(x, out):
    out.copy_(F.max_pool2d(x.permute(0, 3, 1, 2).float(), 3, 2, 1).permute(0, 2, 3, 1).to(out.dtype))
    return out


def rpn_select(head, n_img, h, w, base_anchors, im_w, im_h, pre_nms, post_nms, nms_thresh, min_size=0.0, stride=16,
               out=None, want_anchor=False):
    """head [n, h, w, ld]: [0, A) objectness logits, [A, 5A) deltas (a*4 + c) -> the oracle's proposal selection"""
    a = base_anchors.shape[0]
    boxes, scores, anchor, count = out
    boxes.zero_(), scores.zero_()
    for i in range(n_img):
        hd = head[i].float()
        logits = hd[..., :a].permute(2, 0, 1)[None]
        deltas = hd[..., a:5 * a].permute(2, 0, 1)[None]
        prop, obj = mo.rpn_select(logits, deltas, im_w, im_h, pre_nms, post_nms, nms_thresh, min_size, stride,
                                  cuda_semantics=True)
        k = prop.shape[0]
        boxes[i, :k], scores[i, :k] = prop, obj
        count[i] = k
    return boxes, scores, anchor, count


def roi_align_nhwc(feat, boxes, roi_batch, scale, ph, pw, sampling_ratio, out):
    k = boxes.shape[0]
    b = roi_batch.float().view(-1, 1) if roi_batch is not None else torch.zeros(k, 1)
    pooled = mo.roi_align(feat.permute(0, 3, 1, 2).float().contiguous(), torch.cat([b, boxes.float()], 1), scale, ph, pw,
                          sampling_ratio)                                   # [K, C, ph, pw]
    out.copy_(pooled.permute(0, 2, 3, 1).reshape(k, -1).to(out.dtype))      # bin-major, channel-minor
    return out


# ------------------------------------------------------------------------------------------- FGFA / DFF helpers
def fgfa_pool_image(img, out):
    p = F.avg_pool2d(img.reshape(1, 3, img.shape[-2], img.shape[-1]) / 255, 2, stride=2, ceil_mode=True)[0]
    out.zero_()
    out[..., 0:3] = p.permute(1, 2, 0).to(out.dtype)
    return out


def fgfa_build_pairs(ring, slots, key_pos, pairs):
    hq, wq = ring.shape[1], ring.shape[2]
    pairs.zero_()
    key = ring[int(slots[key_pos])]
    for f in range(slots.numel()):
        pairs[f, 3:3 + hq, 3:3 + wq, 0:3] = key[..., 0:3]
        pairs[f, 3:3 + hq, 3:3 + wq, 4:7] = ring[int(slots[f])][..., 0:3]
    return pairs


def avgpool2_nhwc(x, out):
    y = F.avg_pool2d(x.permute(0, 3, 1, 2).float(), 2, stride=2, ceil_mode=True, count_include_pad=False)
    out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))
    return out


def dff_warp_scale(key_feats, flow, scale, out):
    c = key_feats.shape[2]
    warped = mo.fgfa_warp(key_feats.permute(2, 0, 1)[None].float(), flow[..., :2].permute(2, 0, 1)[None].float())
    out[..., :c] = (warped[0].permute(1, 2, 0) * scale[..., :c].float()).to(out.dtype)
    return out


def fgfa_aggregate(ring, slots, key_pos, flow, out, feat_channels, embed_channels, weights_out=None):
    """warp every cached [feats | embedding] map along its flow, cosine-similarity weights against the key frame's
    warped embedding, soft-max over frames, weighted sum (generalized_rcnn_fgfa.py:45-76, :206-214)"""
    maps = torch.stack([ring[int(s_)] for s_ in slots]).permute(0, 3, 1, 2).float()          # [L, C, h, w]
    warped = mo.fgfa_warp(maps, flow[..., :2].permute(0, 3, 1, 2).float())
    wf, emb = warped[:, :feat_channels], warped[:, feat_channels:feat_channels + embed_channels]
    en = emb / (torch.norm(emb, dim=1, keepdim=True) + 1e-10)
    ec = en[key_pos:key_pos + 1]
    wts = torch.softmax(torch.sum(en * ec, dim=1, keepdim=True), dim=0)
    out[..., :feat_channels] = torch.sum(wts * wf, dim=0).permute(1, 2, 0).to(out.dtype)
    if weights_out is not None:
        weights_out.copy_(wts[:, 0].reshape(weights_out.shape))
    return out


class _Event(object):
    def record(self, *a):
        pass

    def synchronize(self):
        pass


class _Stream(object):
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass


@contextlib.contextmanager
def _stream_ctx(stream):
    yield


@contextlib.contextmanager
def cpu_ops():
    from mega_core import _lib
    from mega_core.b200 import ops
    mine = {"conv_gemm": _shadow_conv_gemm, "gather_rows": gather_rows, "copy_rows": copy_rows, "copy_batch": copy_batch,
            "relation_softmax": relation_softmax, "box_postprocess": box_postprocess, "stem_prep": stem_prep,
            "maxpool3x3s2": maxpool3x3s2, "rpn_select": rpn_select, "roi_align_nhwc": roi_align_nhwc,
            "fgfa_pool_image": fgfa_pool_image, "fgfa_build_pairs": fgfa_build_pairs, "avgpool2_nhwc": avgpool2_nhwc,
            "dff_warp_scale": dff_warp_scale, "fgfa_aggregate": fgfa_aggregate}
    saved_ops = {n: getattr(ops, n) for n in mine}
    saved = (torch.Tensor.pin_memory, torch.cuda.Event, _lib.require_cuda, ops.require_cuda, ops.AUTOTUNE[0])
    saved_streams = (torch.cuda.current_stream, torch.cuda.Stream, torch.cuda.stream)
    torch.cuda.current_stream, torch.cuda.Stream, torch.cuda.stream = (lambda *a, **k: _Stream()), _Stream, _stream_ctx
    saved_chains = ops.CHAINS_ENABLED[0]
    ops.CHAINS_ENABLED[0] = False                   # fp16 engines: per-layer calls instead of the persistent chain kernel
    for n, f in mine.items():
        setattr(ops, n, f)
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.Event = _Event
    _lib.require_cuda = ops.require_cuda = lambda *a: None
    try:
        yield
    finally:
        for n, f in saved_ops.items():
            setattr(ops, n, f)
        torch.Tensor.pin_memory, torch.cuda.Event, _lib.require_cuda, ops.require_cuda, ops.AUTOTUNE[0] = saved
        ops.CHAINS_ENABLED[0] = saved_chains
        torch.cuda.current_stream, torch.cuda.Stream, torch.cuda.stream = saved_streams
