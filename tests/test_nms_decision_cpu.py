"""The greedy RPN NMS kernel (csrc/proposals.cu, iou_plus1_gt) decides `RN(inter / union) > thresh` without dividing
whenever inter is outside a 2^-20 relative band around thresh * union, and divides inside the band. This restates that
decision rule in numpy float32 (round-to-nearest, like the kernel's __f*_rn intrinsics) and checks it against the plain
division on random and on adversarial (within +-64 ulps of the threshold) inputs: the keep lists - int64 indices, the
bit-exact part of the parity bar (nms_cpu.cpp:6-75 / nms.cu:16-19) - cannot change."""
import numpy as np
import pytest

f = np.float32


def _decide(inter, u, t):
    t = f(t)
    t_lo, t_hi = f(t * f(1 - 2.0 ** -20)), f(t * f(1 + 2.0 ** -20))
    with np.errstate(all="ignore"):
        exact = (inter / u) > t
        fast = np.where(u > 0, np.where(inter > (t_hi * u).astype(f), True,
                                        np.where(inter < (t_lo * u).astype(f), False, exact)), exact)
    return exact, fast


@pytest.mark.parametrize("thresh", [0.7, 0.5, 0.3, 0.999, 1e-3])
def test_division_free_iou_decision_is_exact(thresh):
    rng = np.random.default_rng(7)
    u = np.exp(rng.uniform(0, 28, 500_000)).astype(f)
    inter = (u * rng.uniform(0, 1, u.size)).astype(f)
    exact, fast = _decide(inter, u, thresh)
    assert np.array_equal(exact, fast)
    base = (f(thresh) * u).astype(f)
    steps = rng.integers(-64, 65, u.size)
    inter = base.copy()
    for _ in range(64):
        inter = np.where(steps > 0, np.nextafter(inter, f(np.inf)), np.where(steps < 0, np.nextafter(inter, f(-np.inf)), inter))
        steps = steps - np.sign(steps)
    exact, fast = _decide(inter, u, thresh)
    assert np.array_equal(exact, fast)
    # degenerate unions: zero, negative, inf, nan fall through to the division
    u2 = np.array([0, -1, np.inf, np.nan, 1, 1], dtype=f)
    i2 = np.array([0, 1, 1, 1, np.nan, np.inf], dtype=f)
    exact, fast = _decide(i2, u2, thresh)
    assert np.array_equal(exact, fast)
