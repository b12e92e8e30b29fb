"""configs[4] extension: per-frame multipath taps, LTS channel estimate + one-tap ZF equaliser + slicer.  The
reference has no multipath channel; the oracle is the CPU restatement built on the reference's
estimator / equaliser.  Supplied taps reproduce the oracle bit for bit; on-chip Philox taps agree to 1e-5."""
import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("n_sym,n_taps", [(2, 1), (2, 4), (2, 16), (3, 7)])
def test_supplied_taps_bit_exact_chain(ofdm, pkg, port, n_sym, n_taps):
    n_frames = 200
    bits, g = bits_and_noise(40 + n_taps, n_frames, n_sym)
    taps = port.philox_taps(5, 0, n_frames, n_taps)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    frames = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT, with_power=False)
    faded = ofdm.multipath_taps(frames, ofdm.to_dev(taps), n_sym)
    want_faded = port.apply_taps(port.tx_frames(bits, n_sym), taps)
    assert same(faded.cpu().numpy(), want_faded)
    for snr in (5.0, 15.0, 30.0):
        cnt, d = ofdm.awgn_rx_inject(faded, ofdm.to_dev(g), packed, snr, n_sym, pkg.MODE_EXACT, want=("frame_bit_errors", "frame_evm_lin"))
        acc, fe, fv = port.chain_multipath(bits, g, taps, n_sym, snr, per_frame=True)
        assert same(d["frame_bit_errors"].cpu().numpy(), fe)
        assert cnt.bit_errors == acc.bit_errors and cnt.rail_errors == acc.rail_errors
        assert np.allclose(d["frame_evm_lin"].cpu().numpy(), fv, rtol=1e-5)


def test_philox_taps_and_sweep(ofdm, pkg, port):
    n_frames, n_sym, n_taps, seed = 4000, 2, 8, 31
    bits = ofdm.random_bits(seed, 0, n_frames, n_sym)
    frames = ofdm.tx_frames(bits, n_sym, pkg.MODE_FAST, with_power=False)
    faded, taps = ofdm.multipath_philox(frames, seed, 0, n_taps, n_sym)
    want_taps = port.philox_taps(seed, 0, n_frames, n_taps)
    assert np.max(np.abs(taps.cpu().numpy() - want_taps)) < 1e-5      # MUFU log/sin/cos vs libm
    tp = np.sum(taps.cpu().numpy().astype(np.float64) ** 2, axis=(1, 2))
    assert abs(tp.mean() - 1.0) < 0.03                                   # unit expected total power
    snrs = [5.0, 15.0, 25.0, 35.0]
    got = ofdm.mc_sweep_multipath(seed, 0, n_frames, n_sym, n_taps, snrs, pkg.MODE_FAST)
    b = port.philox_bits(seed, 0, n_frames, n_sym)
    prev = 1.0
    for i, s in enumerate(snrs):
        g = port.philox_normals(seed, i, 0, n_frames, 320)
        acc = port.chain_multipath(b, g, want_taps, n_sym, s)
        ber_g, ber_c = got[i].bit_errors / got[i].bits, acc.bit_errors / acc.bits
        assert abs(ber_g - ber_c) <= 0.02 * ber_c + 3.0 / acc.bits, (s, got[i].bit_errors, acc.bit_errors)
        assert ber_g <= prev
        prev = ber_g
    # Rayleigh-like fading: far worse than AWGN at the same SNR, still falling with SNR
    awgn = ofdm.mc_sweep_philox(seed, 0, n_frames, n_sym, [15.0], pkg.MODE_FAST)[0]
    assert got[1].bit_errors > 50 * max(1, awgn.bit_errors)
    # split invariance through the chunked driver
    a = ofdm.mc_sweep_multipath(seed, 0, 1000, n_sym, n_taps, snrs, pkg.MODE_EXACT)
    b0 = ofdm.mc_sweep_multipath(seed, 0, 400, n_sym, n_taps, snrs, pkg.MODE_EXACT)
    b1 = ofdm.mc_sweep_multipath(seed, 400, 600, n_sym, n_taps, snrs, pkg.MODE_EXACT)
    for x, y, z in zip(a, b0, b1):
        assert x.bit_errors == y.bit_errors + z.bit_errors and x.frames == 1000


def test_c_driver_multipath_sweep(tmp_path):
    """configs[4] through the C host driver (EVM vs SNR file): EVM falls with SNR, BER has the fading floor."""
    import os, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "ofdm_sweep")
    if not os.path.exists(exe):
        import __graft_entry__ as entry
        entry.build()
    out = tmp_path / "data"; out.mkdir()
    r = subprocess.run([exe, "--quiet", "--outdir", str(out), "--frames", "100000", "--taps", "8", "--snr-start", "5", "--snr-count", "8",
                        "--snr-step", "5", "--mode", "fast"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
    evm = [float(w) for w in open(out / "Output_EVM_AGC.txt").read().split()]
    assert len(ber) == 8 and all(a >= b for a, b in zip(ber, ber[1:])) and ber[0] > 0.05 and 0 < ber[-1] < ber[0] / 10
    assert all(np.isfinite(e) for e in evm)


@pytest.mark.parametrize("n_taps", [1, 5, 8, 16])
def test_fused_multipath_sweep_equals_staged(ofdm, pkg, n_taps):
    """configs[4] on chip (k_mc_philox<., multipath>) against the staged path (TX, k_multipath, frame power, per-SNR receiver
    kernels over HBM): the same taps and draws -> identical integer totals in EXACT mode, fp32 agreement in FAST mode"""
    snr = [0.0, 6.0, 12.0, 18.0]
    n = 40_000
    try:
        for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
            ofdm.set_option("multipath_path", 2)
            fused = ofdm.mc_sweep_multipath(21, 700, n, 2, n_taps, snr, mode)
            ofdm.set_option("multipath_path", 1)
            staged = ofdm.mc_sweep_multipath(21, 700, n, 2, n_taps, snr, mode)
            ofdm.set_option("force_generic_rx", 1)                      # ... and through the generic receiver kernel
            generic = ofdm.mc_sweep_multipath(21, 700, n, 2, n_taps, snr, mode)
            ofdm.set_option("force_generic_rx", 0)
            if mode == pkg.MODE_EXACT:
                for a, b in zip(generic, staged):
                    assert (a.bit_errors, a.frames_in_error, a.rail_errors) == (b.bit_errors, b.frames_in_error, b.rail_errors), (n_taps, mode)
                for a, b in zip(fused, staged):
                    assert (a.bit_errors, a.bits, a.frames_in_error, a.rail_errors, a.frames) == \
                           (b.bit_errors, b.bits, b.frames_in_error, b.rail_errors, b.frames), (n_taps, mode)
                    assert abs(a.sum_err2 - b.sum_err2) <= 1e-5 * b.sum_err2, (n_taps, mode)
            else:
                # FAST mode: plain fp32 with Philox noise (no EVM guard, statistical results).  The staged route's streaming
                # receiver transforms the sum of the LTS halves once, the fused and the generic kernels each half: rails within
                # fp32 rounding of the slicer boundary may fall either way, and the EVM sum -- dominated by the deepest fades,
                # |e|^2 ~ 1/|H|^2 -- agrees to fp32 accuracy of those few bins, not to 1e-5
                for other in (generic, fused):
                    for a, b in zip(other, staged):
                        assert (a.bits, a.frames) == (b.bits, b.frames)
                        assert abs(a.bit_errors - b.bit_errors) <= 2 + 1e-5 * b.bit_errors and abs(a.rail_errors - b.rail_errors) <= 2 + 1e-5 * b.rail_errors
                        assert abs(a.frames_in_error - b.frames_in_error) <= 2
                        assert abs(a.sum_err2 - b.sum_err2) <= 5e-4 * b.sum_err2, (n_taps, mode)
    finally:
        ofdm.set_option("force_generic_rx", 0)
        ofdm.set_option("multipath_path", 0)
