"""The exact frame power (Transmission_Over_Air, OFDM.c:637-643) at memory speed: k_frame_power_tiled speculates every term
of the serial float chain as x^2 + y^2 in double and takes the reference's operations (glibc hypot, squared) wherever the
running sum comes within `power_margin` double ulps of a tie between two floats.  Whatever the margin, the result must be
the reference's, bit for bit: margin 2^28 sends every sample through the reference's operations, 2^24 about 6 % of them."""
import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def powers(ofdm, pkg, frames, margins=(16, 1 << 24, 1 << 28)):
    out = []
    for m in margins:
        ofdm.set_option("power_margin", m)
        out.append(ofdm.frame_power(frames, pkg.MODE_EXACT).cpu().numpy())
    ofdm.set_option("power_margin", 16)
    return out


@pytest.mark.parametrize("n_sym", [1, 2, 3, 7])
def test_transmitter_power_against_oracle(ofdm, pkg, port, n_sym):
    """frames of the transmitter (LTS prefix constant + tiled chain) and of ofdm_frame_power (tiled chain over the whole
    frame) against the oracle's frame_power, for frame counts that do not fill a warp's 32 rows"""
    for n_frames in (1, 31, 33, 257, 1000):
        bits, _ = bits_and_noise(40 + n_sym + n_frames, n_frames, n_sym)
        packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
        frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
        want = np.array([port.frame_power(f) for f in frames.cpu().numpy()], np.float32)
        assert same(power.cpu().numpy(), want)
        for got in powers(ofdm, pkg, frames):
            assert same(got, want)


def test_margins_agree_on_hard_inputs(ofdm, pkg):
    """random waveforms over 40 binades per frame set, zeros, tiny and huge samples: every margin gives the same floats"""
    import torch
    dev = ofdm.device
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    n = 200_000
    x = torch.randn((n, 320, 2), dtype=torch.float32, device=dev, generator=gen)
    scale = torch.exp2(torch.randint(-20, 20, (n, 1, 1), device=dev, generator=gen).to(torch.float32))
    x *= scale
    x[::7, 5:40] = 0.0                                   # runs of zero samples (term 0: the sum repeats)
    x[::11, 100] = 3.0e-23                               # a negligible sample beside normal ones
    x[1::11, 17, 0] = 3.0e19                             # one sample that dominates the sum (overflows the float range squared)
    x[2::11, :, :] *= 1.0e-22                            # whole frames whose sums live among the float subnormals
    x[3::11, :, 1] = 0.0                                 # real-only frames: hypot(x, 0) = |x| exactly
    a, b, c = powers(ofdm, pkg, x)
    assert same(a, c), np.flatnonzero(a != c)[:10]
    assert same(b, c), np.flatnonzero(b != c)[:10]
    assert np.isinf(c[1::11]).all()
    assert (c[2::11] < 1e-30).all()
    # the per-thread chain (the kernel for lengths that are not whole 16-sample chunks) on the same samples, one sample longer
    y = torch.zeros((4096, 322, 2), dtype=torch.float32, device=dev)
    y[:, :320] = x[:4096]
    d = ofdm.frame_power(y, pkg.MODE_EXACT).cpu().numpy()
    e = (c[:4096].astype(np.float64) * 320 / 322).astype(np.float32)      # same chain, zeros appended: same sum, other divisor
    assert np.allclose(d, e, rtol=2e-7, atol=1e-37)
