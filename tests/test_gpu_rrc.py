"""SURVEY 8(f) rank 1 -- x2 oversampling + RRC pulse shaping around the stage chain, against the reference's own
Convolution() (through oracle/_ref when present, else the port, which the CPU suite pins to it): bit-exact."""
import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("n_sym", [1, 2, 4])
def test_rrc_tx_rx_bit_exact_and_chain(ofdm, pkg, port, po, n_sym):
    oracle = po.Ref() if po.have_ref() else port
    n_frames, L = 150, 160 + 80 * n_sym
    bits, _ = bits_and_noise(70 + n_sym, n_frames, n_sym)
    rng = np.random.default_rng(n_sym)
    g = rng.standard_normal((n_frames, 2 * L + 20)).astype(np.float32)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    frames = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT, with_power=False)
    shaped = ofdm.rrc_tx(frames)
    want_shaped = oracle.rrc_tx(port.tx_frames(bits, n_sym))
    assert same(shaped.cpu().numpy(), want_shaped)
    for snr in (8.0, 25.0):
        ota = ofdm.awgn_inject_len(shaped, ofdm.to_dev(g), snr, pkg.MODE_EXACT)
        want_ota = port.awgn_inject(want_shaped, g, snr)
        assert same(ota.cpu().numpy(), want_ota)
        for idx in (20, 0, 7):
            rx = ofdm.rrc_rx(ota, idx, L)
            assert same(rx.cpu().numpy(), oracle.rrc_rx(want_ota, idx, L))
        # back through the receiver at the aligned index: same decisions as the oracle chain
        rx = ofdm.rrc_rx(ota, 20, L)
        cnt, d = ofdm.rx_frames(rx, packed, n_sym, pkg.MODE_EXACT, want=("frame_bit_errors",))
        want = port.rx_frames(oracle.rrc_rx(want_ota, 20, L), bits, n_sym)
        assert same(d["frame_bit_errors"].cpu().numpy(), want["bit_errors"])
    # shape / bounds errors are reported, not executed (the reference would read past Rx_filter_signal)
    assert ofdm.lib.ofdm_rrc_rx(ofdm.h, shaped.data_ptr(), frames.data_ptr(), n_frames, 2 * L + 20, 100, L) == 1
