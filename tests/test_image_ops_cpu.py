"""CPU checks of the device-side test-time transform (mega_core/data/transforms,
csrc/image_ops.cuh):
  1. the numpy restatement of Pillow's two-pass 8-bit resampler equals PIL.Image.resize bit for bit (so the integer
     arithmetic the kernel implements is understood, not guessed);
  2. the kernel's per-pixel body, compiled for the host and driven by the product's own DeviceTestTransform (coefficient
     tables, size rule, argument order), reproduces the reference's pipeline F.resize -> F.to_tensor -> [[2,1,0]] * 255
     -> F.normalize bit for bit on up-scaling, down-scaling, one-axis and identity cases.
The GPU test (tests/test_zz_train_ops_gpu.py) repeats 2 on the device."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import image_oracle as io  # noqa: E402

NATIVE = os.path.join(ROOT, "tests", "native")
MEAN, STD = [102.9801, 115.9465, 122.7717], [1.0, 1.0, 1.0]          # config/defaults.py:51-55


def _frame(seed, h, w):
    g = np.random.default_rng(seed)
    base = g.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))[:h, :w]           # blocks: edges for the filter to smear
    noise = g.integers(0, 40, (h, w, 3), dtype=np.int16)
    return np.clip(img.astype(np.int16) + noise - 20, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,oh,ow", [(72, 128, 60, 100), (48, 64, 120, 160), (37, 53, 37, 80), (90, 41, 33, 41),
                                       (30, 30, 7, 5)])
def test_pillow_resampler_restatement_is_bit_exact(h, w, oh, ow):
    from PIL import Image
    img = _frame(h * 1000 + w, h, w)
    ref = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(io.resize_restated(img, oh, ow), ref)


@pytest.fixture
def host_transform(monkeypatch):
    from mega_core import _lib
    from mega_core.data import transforms as T
    so = os.path.join(NATIVE, "libimage_ops_host.so")
    srcs = [os.path.join(NATIVE, "image_ops_host.cpp"), os.path.join(ROOT, "mega.pytorch_b200", "csrc", "image_ops.cuh"),
            os.path.join(ROOT, "include", "mega_b200.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-I",
                               os.path.join(ROOT, "mega.pytorch_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                               "-o", so, srcs[0]])
    host = ctypes.CDLL(so)
    fn = host.mega_image_transform_u8
    fn.argtypes, fn.restype = _lib.lib.mega_image_transform_u8.argtypes, _lib.lib.mega_image_transform_u8.restype
    monkeypatch.setattr(_lib.lib, "mega_image_transform_u8", fn)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    return T


@pytest.mark.parametrize("h,w,min_size,max_size", [(72, 128, 60, 100), (90, 160, 60, 100), (48, 64, 96, 160),
                                                   (60, 100, 60, 100), (100, 60, 60, 100), (64, 200, 60, 100)])
def test_device_transform_body_matches_reference_pipeline(host_transform, h, w, min_size, max_size):
    img = _frame(h * 1000 + w, h, w)
    ref = io.reference_pipeline(img, min_size, max_size, MEAN, STD, True)
    tr = host_transform.DeviceTestTransform(min_size, max_size, MEAN, STD, True, device="cpu")
    out, tgt = tr(img)
    assert tgt is None and out.shape == ref.shape
    assert torch.equal(out, ref), (out - ref).abs().max()
    from PIL import Image
    out2, _ = tr(Image.fromarray(img, "RGB"))                     # the reference hands the transform a PIL image
    assert torch.equal(out2, ref)
    planar = torch.from_numpy(img).permute(2, 0, 1).contiguous()  # [3, H, W]: the layout of a GPU JPEG decoder's output
    out3, _ = tr(planar)
    assert torch.equal(out3, ref)


def test_device_transform_rgb_unit_range(host_transform):
    """TO_BGR255 = False (torchvision-style models): RGB order, 0..1 range, ImageNet statistics"""
    img = _frame(7, 50, 70)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    ref = io.reference_pipeline(img, 40, 66, mean, std, False)
    out, _ = host_transform.DeviceTestTransform(40, 66, mean, std, False, device="cpu")(img)
    assert torch.equal(out, ref)


def test_size_rule_and_build_transforms(host_transform):
    assert host_transform.get_size((1280, 720), 600, 1000) == (562, 999)       # 720p: the reference's own rounding (transforms.py:44, :54)
    assert host_transform.get_size((640, 480), 600, 1000) == (600, 800)
    assert host_transform.get_size((1000, 600), 600, 1000) == (600, 1000)

    class _In:
        MIN_SIZE_TEST, MAX_SIZE_TEST, PIXEL_MEAN, PIXEL_STD, TO_BGR255 = 600, 1000, MEAN, STD, True

    class _Cfg:
        INPUT = _In
    from mega_core.data.transforms import build_transforms
    tr = build_transforms(_Cfg, is_train=False, device="cpu")
    assert tr.min_size == 600 and tr.max_size == 1000
    with pytest.raises(NotImplementedError):
        build_transforms(_Cfg, is_train=True)
