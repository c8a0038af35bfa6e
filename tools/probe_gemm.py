"""GPU probe: accuracy of the TF32 path (TMA rounding on/off) and kernel throughput at the
hot-path shapes. Writes gpurun_out/probe_gemm.json. Diagnostic tool, not a test."""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mega.pytorch_b200"))
from mega_core import _lib  # noqa: E402
from mega_core.b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
res = {"device": torch.cuda.get_device_name(0), "device_ok": _lib.lib.mega_device_ok()}


def rel(got, ref):
    ref = ref.double()
    return ((got.double().cpu() - ref).abs().max() / ref.pow(2).mean().sqrt()).item()


# ---- accuracy: rounding on vs off
g = torch.Generator().manual_seed(0)
x = torch.randn(2394, 1024, generator=g)
w = torch.randn(1024, 1024, generator=g) / 32
ref = x.double() @ w.double().t()
for mode in (1, 0):
    _lib.lib.mega_set_tf32_rounding(mode)
    out = torch.empty(2394, 1024, device=dev)
    ops.linear(x.to(dev), w.to(dev), out)
    torch.cuda.synchronize()
    res["linear_relerr_round%d" % mode] = rel(out, ref)
_lib.lib.mega_set_tf32_rounding(1)
torch.backends.cuda.matmul.allow_tf32 = True
res["torch_tf32_relerr"] = rel(x.to(dev) @ w.to(dev).t(), ref)
torch.backends.cuda.matmul.allow_tf32 = False
res["torch_fp32_relerr"] = rel(x.to(dev) @ w.to(dev).t(), ref)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


shapes = [
    # name, n, h, w, cin, cout, ks, dil
    ("res4_1x1_1024_256", 2, 38, 63, 1024, 256, 1, 1),
    ("res4_3x3_256", 2, 38, 63, 256, 256, 3, 1),
    ("res4_1x1_256_1024", 2, 38, 63, 256, 1024, 1, 1),
    ("res5_3x3d2_512", 2, 38, 63, 512, 512, 3, 2),
    ("res5_1x1_512_2048", 2, 38, 63, 512, 2048, 1, 1),
    ("rpn_3x3_1024", 2, 38, 63, 1024, 1024, 3, 1),
    ("res2_3x3_64", 2, 150, 250, 64, 64, 3, 1),
    ("res2_1x1_64_256", 2, 150, 250, 64, 256, 1, 1),
    ("res3_3x3_128", 2, 75, 125, 128, 128, 3, 1),
]
perf = {}
for name, n, h, wd, cin, cout, ks, dil in shapes:
    a = torch.randn(n, h, wd, cin, device=dev)
    wp = torch.randn(ks * ks, cout, cin, device=dev) / (cin * ks * ks) ** 0.5
    out = torch.empty(n, h, wd, cout, device=dev)
    sc = torch.ones(cout, device=dev)
    bi = torch.zeros(cout, device=dev)
    flops = 2.0 * n * h * wd * cin * cout * ks * ks
    entry = {}
    for bn in ops.BLOCK_NS:
        if bn >= 2 * cout and bn > 32:
            continue
        for sk in (0, 1):
            try:
                t = timeit(lambda: ops.conv_gemm(a, wp, out, taps=(ks, ks), dil=dil, pad=dil * (ks - 1) // 2,
                                                 scale=sc, bias=bi, relu=True, block_n=bn, stream_k=sk))
                entry["bn%d_sk%d" % (bn, sk)] = {"us": t * 1e6, "tflops": flops / t / 1e12}
            except Exception as e:  # noqa
                entry["bn%d_sk%d" % (bn, sk)] = {"error": str(e)}
    # cuDNN/cuBLAS reference speed (tf32 allowed) for orientation only
    torch.backends.cudnn.allow_tf32 = True
    xn = a.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wn = wp.view(ks, ks, cout, cin).permute(2, 3, 0, 1).contiguous(memory_format=torch.channels_last)
    t = timeit(lambda: F.conv2d(xn, wn, None, 1, dil * (ks - 1) // 2, dil))
    entry["cudnn_tf32"] = {"us": t * 1e6, "tflops": flops / t / 1e12}
    perf[name] = entry
res["perf"] = perf

# big FC (weight-bandwidth bound): 450 x 100352 -> 1024
m, k, n = 450, 100352, 1024
x = torch.randn(m, k, device=dev)
w = torch.randn(n, k, device=dev) / k ** 0.5
out = torch.empty(m, n, device=dev)
fc = {}
for bn in (128, 256):
    t = timeit(lambda: ops.linear(x, w, out, block_n=bn, stream_k=1), iters=10)
    fc["bn%d" % bn] = {"us": t * 1e6, "weight_GBps": k * n * 4 / t / 1e9, "tflops": 2.0 * m * k * n / t / 1e12}
t = timeit(lambda: torch.mm(x, w.t()), iters=10)
fc["cublas_fp32"] = {"us": t * 1e6}
res["fc0"] = fc

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w") as fh:
    json.dump(res, fh, indent=1)
print(json.dumps(res, indent=1))
