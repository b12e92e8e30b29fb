"""The sweep kernel computes the noise scale (float)sqrt((double)(P / snr)) of OFDM.c:647/:651 as the correctly rounded float
square root: rounding the 53-bit square root to 24 bits cannot differ from rounding the exact root (53 >= 2 * 24 + 2).
Checked here on the values most likely to break it: floats next to squares of floats and of float midpoints, and random ones."""
import numpy as np


def test_double_sqrt_rounded_to_float_is_float_sqrt():
    rng = np.random.default_rng(3)
    r = rng.uniform(1.0, 2.0, 200_000).astype(np.float32)
    mid = (r.astype(np.float64) + np.spacing(r).astype(np.float64) / 2)              # midpoints between adjacent floats
    cand = [rng.uniform(1e-6, 1e6, 400_000).astype(np.float32)]
    for base in (r.astype(np.float64) ** 2, mid ** 2):
        q = base.astype(np.float32)
        cand += [q, np.nextafter(q, np.float32(0)), np.nextafter(q, np.float32(np.inf))]
    q = np.concatenate(cand)
    via_double = np.sqrt(q.astype(np.float64)).astype(np.float32)
    direct = np.sqrt(q)                                                                  # IEEE float sqrt
    assert np.array_equal(via_double, direct)
    # and both are the correctly rounded root: compare against extended precision
    ext = np.sqrt(q.astype(np.longdouble)).astype(np.float32)
    assert np.array_equal(direct, ext)
