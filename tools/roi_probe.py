"""timing probe: fp16 NHWC ROIAlign on two roi-size distributions (per-sample kernel vs MEGA_B200_ROI_SEPARABLE=1)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mega.pytorch_b200"))
from mega_core.b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(5)
feat = torch.randn(2, 38, 63, 2048, generator=g).half().to(dev)
res = {}
for name, lo, hi in (("small", 40.0, 360.0), ("large", 200.0, 900.0)):
    k = 375
    xy = torch.rand(k, 2, generator=g) * torch.tensor([600.0, 300.0])
    wh = torch.rand(k, 2, generator=g) * (hi - lo) + lo
    b = torch.cat([xy, xy + wh], 1)
    b[:, 0::2].clamp_(0, 999)
    b[:, 1::2].clamp_(0, 599)
    boxes = b.to(dev)
    bidx = (torch.arange(k) % 2).int().to(dev)
    out = torch.empty(k, 49 * 2048, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.roi_align_nhwc(feat, boxes, bidx, 1.0 / 16, 7, 7, 0, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.roi_align_nhwc(feat, boxes, bidx, 1.0 / 16, 7, 7, 0, out)
    e1.record()
    e1.synchronize()
    res[name] = (round(e0.elapsed_time(e1) / 20 * 1e3, 1), float(out.float().abs().mean()))
print("ROI_SEPARABLE=%s" % os.environ.get("MEGA_B200_ROI_SEPARABLE", "unset"), res)
