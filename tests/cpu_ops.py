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


def _rows(t, n, row_len):
    return t.reshape(-1, t.shape[-1])[:, :row_len] if t.dim() != 2 else t[:, :row_len]


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


class _Event(object):
    def record(self, *a):
        pass

    def synchronize(self):
        pass


@contextlib.contextmanager
def cpu_ops():
    from mega_core import _lib
    from mega_core.b200 import ops
    saved_ops = {n: getattr(ops, n) for n in ("conv_gemm", "gather_rows", "copy_rows", "copy_batch", "relation_softmax",
                                              "box_postprocess")}
    saved = (torch.Tensor.pin_memory, torch.cuda.Event, _lib.require_cuda, ops.require_cuda, ops.AUTOTUNE[0])
    saved_chains = ops.CHAINS_ENABLED[0]
    ops.CHAINS_ENABLED[0] = False                   # fp16 engines: per-layer calls instead of the persistent chain kernel
    ops.conv_gemm, ops.gather_rows, ops.copy_rows, ops.copy_batch = _shadow_conv_gemm, gather_rows, copy_rows, copy_batch
    ops.relation_softmax, ops.box_postprocess = relation_softmax, box_postprocess
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
