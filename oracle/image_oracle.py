"""image_oracle -- CPU oracle of the test-time input transform. TEST INFRASTRUCTURE ONLY.

`reference_pipeline` IS the reference's arithmetic: the same three library calls its transforms make
(data/transforms/transforms.py:59 F.resize on a PIL image, :119 F.to_tensor, :131-133 `image[[2, 1, 0]] * 255` and
F.normalize), executed by the Pillow / torchvision installed here (Pillow 12.2.0, unpinned by the reference).
`resize_restated` is a numpy restatement of Pillow's two-pass 8-bit resampler (src/libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) used to explain and
cross-check the integer arithmetic the CUDA kernel implements; tests/test_image_ops_cpu.py pins it to PIL bit for bit."""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def reference_pipeline(rgb_u8, min_size, max_size, mean, std, to_bgr255=True):
    """uint8 [H,W,3] RGB -> float32 [3,H',W'] exactly as data/transforms/build.py:5-49 does at test time"""
    from PIL import Image
    from torchvision.transforms import functional as F
    img = Image.fromarray(np.ascontiguousarray(rgb_u8), "RGB")
    w, h = img.size
    size = min_size
    if max_size is not None:                                  # Resize.get_size, transforms.py:36-56
        lo, hi = float(min((w, h))), float(max((w, h)))
        if hi / lo * size > max_size:
            size = int(round(max_size * lo / hi))
    if (w <= h and w == size) or (h <= w and h == size):
        oh, ow = h, w
    elif w < h:
        ow, oh = size, int(size * h / w)
    else:
        oh, ow = size, int(size * w / h)
    img = F.resize(img, (oh, ow))
    t = F.to_tensor(img)
    if to_bgr255:
        t = t[[2, 1, 0]] * 255
    return F.normalize(t, mean=list(mean), std=list(std))


def _coeffs(in_size, out_size):
    scale = filterscale = float(in_size) / out_size
    filterscale = max(filterscale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    rows = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)]
        ww = sum(w[:0], 0.0)
        for v in w:
            ww += v
        k = [int(0.5 + (v / ww if ww != 0.0 else v) * (1 << PRECISION_BITS)) for v in w]
        rows.append((xmin, np.asarray(k, dtype=np.int64)))
    return rows


def _pass(img, rows, axis):
    """one 8-bit pass along `axis` (1: horizontal, 0: vertical) of an [H, W, 3] uint8 image"""
    img = img.astype(np.int64)
    out = []
    for first, k in rows:
        seg = img[:, first:first + len(k)] if axis == 1 else img[first:first + len(k)]
        kk = k.reshape(1, -1, 1) if axis == 1 else k.reshape(-1, 1, 1)
        acc = (1 << (PRECISION_BITS - 1)) + (seg * kk).sum(axis=axis)
        out.append(np.clip(acc >> PRECISION_BITS, 0, 255))
    return np.stack(out, axis=axis).astype(np.uint8)


def resize_restated(rgb_u8, oh, ow):
    """PIL.Image.resize((ow, oh), BILINEAR) of an RGB image, restated: horizontal pass, round to uint8, vertical pass"""
    h, w = rgb_u8.shape[:2]
    img = rgb_u8
    if ow != w:
        img = _pass(img, _coeffs(w, ow), 1)
    if oh != h:
        img = _pass(img, _coeffs(h, oh), 0)
    return img
