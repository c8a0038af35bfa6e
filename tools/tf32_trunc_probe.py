"""Does tcgen05 kind::tf32 truncate or round the low 13 mantissa bits of a raw fp32 operand? (TMA conversion off, so the MMA
sees the fp32 bit pattern.)  x = 1 + 0.75 * 2^-10: truncation gives 1.0, round-to-nearest 1 + 2^-10."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mega.pytorch_b200"))
import torch  # noqa: E402

from mega_core._lib import lib  # noqa: E402
from mega_core.b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
ops.AUTOTUNE[0] = False
old = lib.mega_set_tf32_rounding(0)
for val in (1.0 + 0.75 * 2 ** -10, 1.0 + 0.25 * 2 ** -10, -(1.0 + 0.75 * 2 ** -10)):
    x = torch.full((128, 32), 0.0, device=dev)
    x[:, 0] = val                                   # A row = [val, 0, 0, ...]
    w = torch.zeros(32, 32, device=dev)
    w[0, 0] = 1.0                                   # out[:, 0] = A[:, 0] * 1
    w[1, 0] = val                                   # out[:, 1] = A[:, 0] * val  (B side)
    out = torch.zeros(128, 32, device=dev)
    with ops.precision("tf32"):
        ops.linear(x, w, out, block_n=32, stream_k=0)
    torch.cuda.synchronize()
    print("x = %.10f -> A-side product with 1.0: %.10f ; A * B(val): %.10f  (trunc: %.10f, RN: %.10f)" % (
        val, out[0, 0].item(), out[0, 1].item(), 1.0 if val > 0 else -1.0,
        (1 + 2 ** -10) * (1 if val > 0 else -1) if abs(val) > 1 + 0.5 * 2 ** -10 else (1.0 if val > 0 else -1.0)))
lib.mega_set_tf32_rounding(old)
