"""The receiver's stages one by one (ofdm_strip_cp, ofdm_fft64, ofdm_channel_estimate, ofdm_equalize, ofdm_demap,
ofdm_agc_slicer, ofdm_qpsk_demodulate): the 1:1 batched counterparts of the reference's Channel_Estimation (src/OFDM.c:830),
CP strip (:1024), fft (:314), equaliser (:1046), demap (:1061), AGC_Receiver (:852) and QPSK_Demodulator (:873).  Chained,
they must reproduce the golden vectors generated from the compiled reference (tests/golden/make_golden.py) bit for bit."""
import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("n_sym", [2, 5])
def test_stage_chain_reproduces_the_reference_vectors(ofdm, pkg, golden, n_sym):
    g, t = golden, "n%d_" % n_sym
    n_frames = g[t + "bits"].shape[0]
    for i, snr in enumerate(g[t + "snr"]):
        ota = ofdm.to_dev(g[t + "ota_%d" % i])
        H = ofdm.channel_estimate(ota, pkg.MODE_EXACT)                                   # :830
        assert same(H.cpu().numpy(), g[t + "rx_H_%d" % i])
        bodies = ofdm.strip_cp(ota, n_sym)                                               # :1024
        assert same(bodies.cpu().numpy(), g[t + "ota_%d" % i][:, 160:].reshape(n_frames, n_sym, 80, 2)[:, :, 16:])
        F = ofdm.fft64(bodies, pkg.MODE_EXACT).reshape(n_frames, n_sym, 64, 2)           # :1037
        E = ofdm.equalize(F, H, pkg.MODE_EXACT)                                          # :1046
        pts = ofdm.demap(E)                                                              # :1061
        assert same(pts.cpu().numpy().reshape(n_frames, n_sym * 48, 2), g[t + "rx_eq_%d" % i])
        sl = ofdm.agc_slicer(pts)                                                        # :852
        assert same(sl.cpu().numpy().reshape(n_frames, n_sym * 48, 2), g[t + "rx_sliced_%d" % i])
        bits = ofdm.qpsk_demodulate(sl)                                                  # :873
        assert same(pkg.unpack_bits_host(bits.cpu().numpy().view(np.uint32)).reshape(n_frames, -1), g[t + "rx_bits_%d" % i])
        # the null bins of the equaliser output hold what the reference's division leaves there: nothing finite and usable
        e = E.cpu().numpy()
        nulls = [0, 1, 2, 3, 4, 5, 32, 59, 60, 61, 62, 63]
        assert not np.isfinite(e[:, :, nulls, :]).all()


def test_stages_against_the_oracle_on_random_and_degenerate_frames(ofdm, pkg, port):
    n_frames, n_sym = 300, 3
    bits, g = bits_and_noise(99, n_frames, n_sym)
    ota = port.awgn_inject(port.tx_frames(bits, n_sym), g, 2.0)
    ota[0] = 0.0                                         # H = 0: the division's recovery branch
    ota[1, 32:96] = -ota[1, 96:160]                      # cancelling LTS halves: signed zeros in H
    ota[2] *= np.float32(1e-25)
    ota[3] *= np.float32(1e15)
    want = port.rx_frames(ota, bits, n_sym)
    od = ofdm.to_dev(ota)
    H = ofdm.channel_estimate(od, pkg.MODE_EXACT)
    assert same(H.cpu().numpy(), want["H"])
    assert np.array_equal(np.signbit(H.cpu().numpy()), np.signbit(want["H"]))            # even the signs of zeros
    F = ofdm.fft64(ofdm.strip_cp(od, n_sym), pkg.MODE_EXACT).reshape(n_frames, n_sym, 64, 2)
    pts = ofdm.demap(ofdm.equalize(F, H, pkg.MODE_EXACT))
    assert same(pts.cpu().numpy().reshape(n_frames, -1, 2), want["eq"])
    sl = ofdm.agc_slicer(pts)
    assert same(sl.cpu().numpy().reshape(n_frames, -1, 2), want["sliced"])
    bits_rx = pkg.unpack_bits_host(ofdm.qpsk_demodulate(sl).cpu().numpy().view(np.uint32)).reshape(n_frames, -1)
    assert same(bits_rx, want["bits"])
    # fast mode: same stages in fp32, 1e-5 of each frame's peak
    Hf = ofdm.channel_estimate(od, pkg.MODE_FAST).cpu().numpy()
    ok = [i for i in range(n_frames) if i not in (0, 1)]
    scale = np.abs(want["H"][ok]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(Hf[ok] - want["H"][ok]) / scale) <= 1e-5


def test_stage_frames_with_an_sts_slot_and_raw_demodulator_inputs(ofdm, pkg, port):
    """the reference's own frame layout STS || LTS || data (Channel_Estimation reads samples 192..319, the CP strip starts at 320),
    and QPSK_Demodulator's own comparisons on points that did not come out of the slicer (zeros and NaNs go to 11)"""
    n_frames, n_sym = 50, 2
    bits, g = bits_and_noise(5, n_frames, n_sym)
    ota = port.awgn_inject(port.tx_frames(bits, n_sym), g, 6.0)
    full = ofdm.prepend_sts(ofdm.to_dev(ota))
    H0 = ofdm.channel_estimate(ofdm.to_dev(ota), pkg.MODE_EXACT)
    H1 = ofdm.channel_estimate(full, pkg.MODE_EXACT, lts_off=160)
    assert same(H0.cpu().numpy(), H1.cpu().numpy())
    assert same(ofdm.strip_cp(full, n_sym, data_off=320).cpu().numpy(), ofdm.strip_cp(ofdm.to_dev(ota), n_sym).cpu().numpy())
    pts = np.zeros((1, 48, 2), np.float32)
    pts[0, 0] = (1, 1); pts[0, 1] = (-1, 1); pts[0, 2] = (-1, -1); pts[0, 3] = (1, -1)
    pts[0, 4] = (0, 1); pts[0, 5] = (np.nan, 1); pts[0, 6] = (-1, 0); pts[0, 7] = (0, 0)
    got = pkg.unpack_bits_host(ofdm.qpsk_demodulate(ofdm.to_dev(pts)).cpu().numpy().view(np.uint32)).reshape(-1)
    assert list(got[:16]) == [0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]
    sl = ofdm.agc_slicer(ofdm.to_dev(pts)).cpu().numpy()
    q = np.float32(1 / np.sqrt(2.0))
    assert sl[0, 4, 0] == -q and sl[0, 5, 0] == -q and sl[0, 7, 1] == -q and sl[0, 0, 0] == q      # 0 and NaN go negative (:860-868)


def test_stage_entry_points_reject_bad_arguments(ofdm, pkg):
    lib, h = ofdm.lib, ofdm.h
    x = ofdm.zeros((4, 320, 2), ofdm.torch.float32)
    assert lib.ofdm_strip_cp(h, x.data_ptr(), None, 4, 2, 320, 160) == 1
    assert lib.ofdm_strip_cp(h, x.data_ptr(), x.data_ptr(), 4, 2, 200, 160) == 1          # frame too short for two symbols
    assert lib.ofdm_channel_estimate(h, x.data_ptr(), None, 4, 320, 0, 0) == 1
    assert lib.ofdm_channel_estimate(h, x.data_ptr(), x.data_ptr(), 4, 100, 0, 0) == 1
    assert lib.ofdm_equalize(h, None, None, None, 4, 2, 0) == 1
    assert lib.ofdm_demap(h, None, None, 1) == 1 and lib.ofdm_agc_slicer(h, None, None, 1) == 1 and lib.ofdm_qpsk_demodulate(h, None, None, 1) == 1
    for fn in (lib.ofdm_demap, lib.ofdm_agc_slicer, lib.ofdm_qpsk_demodulate):
        assert fn(h, None, None, 0) == 0                                                   # empty batch is fine
