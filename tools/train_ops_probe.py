"""Times the training-side kernels of ABI v4 and the device image transform at detection-sized workloads (CUDA events,
3 warm-ups, median of 10): the first measurements these kernels get (round 2). Prints achieved GB/s against the
algorithmic bytes, the roofline that bounds them."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mega.pytorch_b200"))
from mega_core import _C  # noqa: E402
from mega_core.data.transforms import DeviceTestTransform  # noqa: E402


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


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    # ROIAlign backward: 512 rois x 1024 channels x 7x7 on a 38x63 map (the res4 map of a 600x1000 frame)
    k, c, h, w = 512, 1024, 38, 63
    x1 = torch.rand(k, generator=g) * 800
    y1 = torch.rand(k, generator=g) * 450
    rois = torch.stack([torch.zeros(k), x1, y1, x1 + 30 + torch.rand(k, generator=g) * 300,
                        y1 + 30 + torch.rand(k, generator=g) * 200], 1).to(dev)
    grad = torch.randn(k, c, 7, 7, device=dev)
    ms = timed(lambda: _C.roi_align_backward(grad, rois, 1 / 16.0, 7, 7, 1, c, h, w, 0))
    print("roi_align_backward  %7.3f ms   grad read %.0f MB -> %.0f GB/s (scatter-bound)" % (ms, grad.numel() * 4 / 1e6,
                                                                                          grad.numel() * 4 / ms / 1e6))
    feat = torch.randn(1, c, h, w, device=dev)
    ms = timed(lambda: _C.roi_pool_forward(feat, rois, 1 / 16.0, 7, 7))
    print("roi_pool_forward    %7.3f ms" % ms)
    # modulated deformable conv backward: 256 -> 256 channels, 3x3, 50 x 84 map, batch 2 (a res4 DCN block)
    b, ci, co, hh, ww = 2, 256, 256, 50, 84
    x = torch.randn(b, ci, hh, ww, device=dev)
    wt = torch.randn(co, ci, 3, 3, device=dev) * 0.02
    off = torch.randn(b, 18, hh, ww, device=dev)
    msk = torch.rand(b, 9, hh, ww, device=dev)
    go = torch.randn(b, co, hh, ww, device=dev)
    bias = torch.zeros(co, device=dev)

    def dcn():
        gi, gw, gb = torch.zeros_like(x), torch.zeros_like(wt), torch.zeros_like(bias)
        goff, gm = torch.zeros_like(off), torch.zeros_like(msk)
        _C.modulated_deform_conv_backward(x, wt, bias, None, off, msk, None, gi, gw, gb, goff, gm, go, 3, 3, 1, 1, 1, 1, 1, 1,
                                          1, 1, True)
    ms = timed(dcn)
    flops = 2 * 2.0 * b * hh * ww * co * ci * 9
    print("modulated_dcn_bwd   %7.3f ms   %.1f GFLOP in the two GEMMs -> %.1f TFLOP/s incl. im2col / col2im" % (ms, flops / 1e9, flops / ms / 1e9))
    # input transform: 720p frame -> 562 x 999
    tr = DeviceTestTransform(600, 1000, [102.9801, 115.9465, 122.7717], [1.0, 1.0, 1.0], True, device=dev)
    img = torch.randint(0, 256, (720, 1280, 3), dtype=torch.uint8, device=dev)
    out = torch.empty(3, 562, 999, device=dev)
    ms = timed(lambda: tr(img, out=out))
    byt = img.numel() + out.numel() * 4
    print("image_transform_u8  %7.3f ms   %.1f MB in + out -> %.0f GB/s" % (ms, byt / 1e6, byt / ms / 1e6))


if __name__ == "__main__":
    main()
