"""Frames far longer than the default (up to hundreds of symbols): the FIR-type kernels (RRC shaping, matched filter,
multipath taps) stage a frame through shared memory in tiles, so their footprint does not grow with the frame; and the C
driver's whole over-the-air path decodes messages of any length, as the reference does (src/OFDM.c:435-465)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("n_sym", [30, 57, 200])
def test_fir_kernels_on_long_frames(ofdm, pkg, port, po, n_sym):
    oracle = po.Ref() if po.have_ref() else port
    n_frames, L = 9, 160 + 80 * n_sym                       # 200 symbols: 16160 samples, 32340 after x2 + RRC (16 tiles)
    bits, _ = bits_and_noise(31 + n_sym, n_frames, n_sym)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    frames = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT, with_power=False)
    want_tx = port.tx_frames(bits, n_sym)
    assert same(frames.cpu().numpy(), want_tx)
    shaped = ofdm.rrc_tx(frames)
    want_shaped = oracle.rrc_tx(want_tx)
    assert same(shaped.cpu().numpy(), want_shaped)
    for idx in (20, 3):
        assert same(ofdm.rrc_rx(shaped, idx, L).cpu().numpy(), oracle.rrc_rx(want_shaped, idx, L))
    # per-capture packet indices, including a packet that starts so late that the filter runs off the capture
    idx = np.array([20, 0, 21, 400, 2 * L, 2 * L + 19, 5, 20, 20], np.int32)
    got = ofdm.rrc_rx_idx(shaped, ofdm.to_dev(idx), L).cpu().numpy()
    for f in (0, 1, 2, 6):
        assert same(got[f], oracle.rrc_rx(want_shaped[f:f + 1], int(idx[f]), L)[0])
    assert np.all(got[5, 11:] == 0) and np.all(got[4, 20:] == 0)          # beyond the filtered capture: zeros, not out-of-bounds reads
    # multipath taps on the long frame == the oracle's convolution, bit for bit
    rng = np.random.default_rng(n_sym)
    taps = (rng.standard_normal((n_frames, 16, 2)) * 0.25).astype(np.float32)
    faded = ofdm.multipath_taps(frames, ofdm.to_dev(taps), n_sym)
    assert same(faded.cpu().numpy(), port.apply_taps(want_tx, taps))


def test_c_driver_long_message_whole_path(tmp_path):
    """a 4-symbol and a 40-symbol message through the reference's whole over-the-air path (the default mode of the driver)"""
    exe = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "ofdm_sweep")
    out = tmp_path / "data"
    out.mkdir()
    for msg in ("The quick brown fox jumps over the lazy dog.!", "802.11a OFDM QPSK " * 26):
        assert len(msg) * 8 > 3 * 96
        r = subprocess.run([exe, "--message", msg, "--snr-start", "30", "--snr-count", "6", "--snr-step", "2", "--outdir", str(out)],
                           capture_output=True, timeout=300)
        stdout = r.stdout.decode("latin-1")
        assert r.returncode == 0, stdout + r.stderr.decode("latin-1")
        assert stdout.count("Received Message: \n" + msg) >= 4               # a late packet in the capture window may garble a point
        ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
        assert len(ber) == 6 and sum(b == 0.0 for b in ber) >= 4


def test_c_driver_until_rule_and_draw_file(tmp_path, ofdm, pkg, port):
    exe = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "ofdm_sweep")
    out = tmp_path / "data"
    out.mkdir()
    # configs[3] as stated: until >= 100 errors or the bit budget, in rounds
    r = subprocess.run([exe, "--quiet", "--outdir", str(out), "--target-errors", "100", "--max-bits", "4000000", "--round-frames", "4096",
                        "--snr-start", "0", "--snr-count", "5", "--snr-step", "4", "--seed", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "until-sweep" in r.stderr
    got, rounds = ofdm.mc_sweep_until(5, 0, 2, 0, [0.0, 4.0, 8.0, 12.0, 16.0], pkg.MODE_EXACT, 100, 4_000_000, 4096)
    ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
    assert len(ber) == 5 and all(abs(b - c.bit_errors / c.bits) <= 5e-3 * max(b, 1e-12) + 1e-12 for b, c in zip(ber, got))    # %.2e text
    assert got[0].frames == 4096 and got[-1].bits >= 4_000_000
    # configs[1] from C: draws from a file (here: the golden reference-captured stream would do; any float32 file works)
    n = 3000
    bits, g = bits_and_noise(11, n, 2)
    packed = pkg.pack_bits_host(bits)
    g.tofile(tmp_path / "draws.f32"); packed.astype(np.uint32).tofile(tmp_path / "bits.u32")
    r = subprocess.run([exe, "--quiet", "--outdir", str(out), "--draws", str(tmp_path / "draws.f32"), "--bits", str(tmp_path / "bits.u32"),
                        "--snr-start", "2", "--snr-count", "4", "--snr-step", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    want = port.chain_sweep(bits, g, 2, [2.0, 5.0, 8.0, 11.0])
    ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
    for b, w in zip(ber, want):
        assert abs(b - w.bit_errors / w.bits) <= 5e-3 * w.bit_errors / w.bits + 1e-12
