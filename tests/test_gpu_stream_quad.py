"""k_stream_quad (csrc/ofdm_stream.cuh): the streaming channel + receiver with one frame per 8-lane group -- a warp works on
four consecutive frames, the two LTS halves are added in time and transformed once, the estimate and the decisions stay in
the owning lane's registers (OFDM.c:830-850, 1018-1165).  Its totals must be those of the CPU oracle and of the
one-frame-per-warp kernels for every batch size (ragged last quad, fewer frames than a quad), frame length, noise source
and block shape."""
import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu


def ints(c):
    return (c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames)


@pytest.fixture()
def knobs(ofdm):
    yield ofdm
    ofdm.set_option("stream_layout", 0)
    ofdm.set_option("stream_warps", 6)
    ofdm.set_option("exact_speculation", 1)
    ofdm.set_option("force_replay", 0)


@pytest.mark.parametrize("n_sym", [1, 2, 3, 7])
def test_ragged_batches_against_the_oracle(knobs, pkg, port, n_sym):
    """1 ... 9 frames and a batch that leaves a partial quad in the last warp: EXACT totals bit-identical to the oracle,
    whatever the layout (one frame per lane group / per warp) and block shape (6 / 8 warps)"""
    ofdm = knobs
    for n_frames in (1, 2, 3, 4, 5, 7, 9, 1187):
        bits, g = bits_and_noise(900 + n_frames + n_sym, n_frames, n_sym)
        packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
        gd = ofdm.to_dev(g)
        frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
        for snr in (1.0, 11.0):
            acc = port.chain(bits, g, n_sym, snr)
            want = (acc.bit_errors, acc.bits, acc.frames_in_error, acc.rail_errors, acc.frames)
            for layout, warps in ((0, 6), (0, 8), (1, 8)):
                ofdm.set_option("stream_layout", layout)
                ofdm.set_option("stream_warps", warps)
                c = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_EXACT, power=power)[0]
                assert ints(c) == want, (n_frames, n_sym, snr, layout, warps)
                assert abs(c.sum_err2 - acc.sum_err2) <= 1e-5 * acc.sum_err2 and abs(c.sum_evm_lin - acc.sum_evm_lin) <= 1e-5 * acc.sum_evm_lin
                f = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_FAST, power=power)[0]
                assert (f.bits, f.frames) == (c.bits, c.frames) and abs(f.bit_errors - c.bit_errors) <= 2
                assert abs(f.sum_err2 - acc.sum_err2) <= 1e-5 * acc.sum_err2


def test_noise_free_and_forced_replay(knobs, pkg, port):
    """noise-free round trip (no errors, any tail) and every frame through the exact replay (same totals)"""
    ofdm = knobs
    n_sym = 2
    for n_frames in (3, 4098):
        bits, g = bits_and_noise(31 + n_frames, n_frames, n_sym)
        packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
        frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
        for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
            c = ofdm.rx_frames(frames, packed, n_sym, mode)[0]
            assert (c.bit_errors, c.frames_in_error, c.rail_errors, c.frames, c.bits) == (0, 0, 0, n_frames, n_frames * 96 * n_sym)
        gd = ofdm.to_dev(g)
        a = ofdm.awgn_rx_inject(frames, gd, packed, 3.0, n_sym, pkg.MODE_EXACT, power=power)[0]
        ofdm.set_option("force_replay", 1)
        ofdm.replayed_frames(reset=True)
        b = ofdm.awgn_rx_inject(frames, gd, packed, 3.0, n_sym, pkg.MODE_EXACT, power=power)[0]
        assert ofdm.replayed_frames() == n_frames
        ofdm.set_option("force_replay", 0)
        acc = port.chain(bits, g, n_sym, 3.0)
        assert ints(a) == ints(b) == (acc.bit_errors, acc.bits, acc.frames_in_error, acc.rail_errors, acc.frames)
        assert abs(b.sum_err2 - acc.sum_err2) <= 1e-5 * acc.sum_err2


def test_philox_noise_matches_the_one_frame_per_warp_kernels(knobs, pkg):
    """on-chip Philox noise (the staged Monte-Carlo / multipath route): EXACT totals identical in both layouts"""
    ofdm = knobs
    import torch
    n, n_sym = 20_001, 2
    gen = torch.Generator(device=ofdm.device); gen.manual_seed(5)
    packed = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * 3 * n_sym,), dtype=torch.int32, device=ofdm.device, generator=gen)
    frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
    for snr in (2.0, 9.0):
        res = []
        for layout in (0, 1):
            ofdm.set_option("stream_layout", layout)
            res.append(ofdm.awgn_rx_philox(frames, packed, snr, 77, 3, 1000, n_sym, pkg.MODE_EXACT, power=power)[0])
        assert ints(res[0]) == ints(res[1]) and res[0].bit_errors > 0
        assert abs(res[0].sum_err2 - res[1].sum_err2) <= 1e-6 * res[1].sum_err2
