"""Timing of the deformable-conv im2col (tile kernel), the modulated variant and the sigmoid focal loss at detection-sized
workloads (CUDA events, median of 10) with the achieved GB/s against their algorithmic bytes -- these kernels are
HBM-bound. `ncu --set full -k regex:"deform_im2col|focal"` on this script gives the committed
captures (profiles/r02_ncu_dcn_*).     python tools/dcn_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mega.pytorch_b200"))
from mega_core import _C, _lib  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
b, c, h, w, k = 2, 256, 50, 84, 3            # a res4-sized map of a 800x1344 frame, 3x3 DCN
x = torch.randn(b, c, h, w, generator=g).to(dev)
off = (torch.randn(b, 2 * k * k, h, w, generator=g) * 2).to(dev)
mask = torch.rand(b, k * k, h, w, generator=g).to(dev)
cols = torch.empty(b, h * w, c * k * k, device=dev)
for name, m in (("deform_im2col (DCN v1)", None), ("deform_im2col (DCN v2, modulated)", mask)):
    def run():
        _lib.check(_lib.lib.mega_deform_im2col(_lib.ptr(x), _lib.ptr(off), _lib.ptr(m) if m is not None else None, b, c, h, w,
                                               k, k, 1, 1, 1, 1, 1, 1, 1, c * k * k, _lib.ptr(cols), _lib.stream_ptr()),
                   "mega_deform_im2col")
    ms = timed(run)
    nbytes = x.numel() * 4 + off.numel() * 4 + cols.numel() * 4 + (m.numel() * 4 if m is not None else 0)
    print("%-36s %7.1f us   %6.1f MB algorithmic -> %6.0f GB/s" % (name, ms * 1e3, nbytes / 1e6, nbytes / ms / 1e6))
n, nc = 120000, 80                            # RetinaNet-sized: anchors x classes
logits = torch.randn(n, nc, generator=g).to(dev)
targets = torch.randint(-1, nc + 1, (n,), generator=g, dtype=torch.int32).to(dev)
ms = timed(lambda: _C.sigmoid_focalloss_forward(logits, targets, nc, 2.0, 0.25))
print("%-36s %7.1f us   %6.1f MB algorithmic -> %6.0f GB/s (incl. the output allocation)" % (
    "sigmoid_focalloss_forward", ms * 1e3, 2 * logits.numel() * 4 / 1e6, 2 * logits.numel() * 4 / ms / 1e6))
d = torch.ones(n, nc, device=dev)
ms = timed(lambda: _C.sigmoid_focalloss_backward(logits, targets, d, nc, 2.0, 0.25))
print("%-36s %7.1f us   %6.1f MB algorithmic -> %6.0f GB/s" % ("sigmoid_focalloss_backward", ms * 1e3,
                                                               3 * logits.numel() * 4 / 1e6, 3 * logits.numel() * 4 / ms / 1e6))
