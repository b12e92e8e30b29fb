"""SURVEY 8(f) rank 2 -- packet detection / selection against the reference's own Packet_Detection and
Packet_Selection (oracle/_ref when present, else the port pinned to it): bit-exact correlation and identical indices,
on captures of the reference's own transmitted waveform (fixture) with noise, plus degenerate captures."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_detection_and_selection_match_reference(ofdm, pkg, port, po, golden):
    oracle = po.Ref() if po.have_ref() else port
    tx = golden["ref_tx_waveform"]                       # Transmitter() output of the reference, 9800 samples
    assert tx.shape == (9800, 2)
    rng = np.random.default_rng(12)
    caps = []
    for trial in range(48):
        snr_db = [2.0, 6.0, 10.0, 20.0][trial % 4]
        p = np.mean(np.sum(tx.astype(np.float64) ** 2, axis=1))
        noisy = tx.copy()
        noisy[:, 0] += (rng.standard_normal(9800) * np.sqrt(p / 10 ** (snr_db / 10))).astype(np.float32)   # real-rail noise (Q1)
        start = int(rng.integers(0, 9800 - 3008))
        caps.append(noisy[start:start + 3008])
    caps.append(np.zeros((3008, 2), np.float32))                                   # silence: 0/0 -> NaN -> no detection
    caps.append(rng.standard_normal((3008, 2)).astype(np.float32))                 # noise only
    # magnitudes that leave the scaled power chain's range (tiny, huge, non-finite): the plain chain must take over
    for scale in (1e-25, 3e-19, 1e17, 1e20):
        caps.insert(0, (caps[len(caps) % 7] * np.float32(scale)).astype(np.float32))
    bad = caps[5].copy(); bad[1000, 0] = np.nan; caps.insert(0, bad)
    bad = caps[6].copy(); bad[2000, 1] = np.inf; caps.insert(0, bad)
    mixed = caps[7].copy(); mixed[::3] *= np.float32(1e-24); caps.insert(0, mixed)
    caps = np.stack(caps)
    corr = ofdm.packet_detect(ofdm.to_dev(caps))
    idx = ofdm.packet_select(corr).cpu().numpy()
    got = corr.cpu().numpy()
    n_found = 0
    for i, cap in enumerate(caps):
        want = oracle.packet_detection(cap)
        assert np.array_equal(got[i], want[:, 0], equal_nan=True), i
        want_idx = oracle.packet_selection(want)
        assert idx[i] == want_idx, (i, idx[i], want_idx)
        n_found += want_idx > 0
    assert n_found >= 20 and idx[-1] == 0 and idx[-2] == 0
