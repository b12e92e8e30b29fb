"""include/ofdm_ref_compat.h: the reference's own function names on host pointers.  A reference-style call site
(tests/compat_callsite.c) must compile and link against it (CPU check), and on a GPU its outputs must be the compiled
reference's, bit for bit -- including Transmission_Over_Air on the same libc rand() stream."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200")


def build_callsite(tmp_path):
    exe = str(tmp_path / "compat_callsite")
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "compat_callsite.c"), "-L", PKG, "-lofdm_b200", "-Wl,-rpath," + PKG, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_style_call_site_compiles_and_links(tmp_path, lib):
    exe = build_callsite(tmp_path)
    out = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    for sym in ("ofdm_qpsk_modulate", "ofdm_fft64", "ofdm_ifft64", "ofdm_awgn_inject_len", "ofdm_channel_estimate", "ofdm_agc_slicer",
                "ofdm_qpsk_demodulate"):
        assert sym in out, sym


@pytest.mark.gpu
def test_reference_style_call_site_matches_the_reference(tmp_path, ofdm, ref, port):
    exe = build_callsite(tmp_path)
    seed, snr = 4321, 7.0
    path = str(tmp_path / "out.bin")
    r = subprocess.run([exe, path, str(seed), str(snr)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "compat call site ok" in r.stdout, r.stdout + r.stderr
    raw = np.fromfile(path, dtype=np.complex64)
    pos = [0]

    def take(n):
        a = raw[pos[0]:pos[0] + n]; pos[0] += n
        return a

    def iq(z):
        return np.stack([z.real, z.imag], axis=-1).astype(np.float32)

    same = lambda a, b: np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
    data = take(192).real.astype(np.uint8).reshape(2, 96)
    mod = take(96)
    assert same(iq(mod).reshape(2, 48, 2), ref.qpsk_mod(data))                                        # QPSK_Modulator :415
    X, Y, Z, X_after = take(64), take(64), take(64), take(64)
    assert same(iq(Y), ref.fft64(iq(X))[0])                                                             # fft :314
    assert same(iq(Z), ref.ifft64(iq(X))[0]) and same(X_after, np.roll(X, 32))                          # ifft :320 (+ its in-place shift)
    tx, ota, H = take(480), take(480), take(64)
    assert same(iq(ota), ref.awgn(iq(tx), snr, seed=seed + 1))                                          # Transmission_Over_Air on the same rand() stream
    assert np.array_equal(ota.imag, tx.imag)                                                            # real-rail noise only (Q1)
    want_H = port.rx_frames(iq(ota)[160:][None], np.zeros((1, 192), np.uint8), 2)["H"][0]               # Channel_Estimation :830 (samples 192..319)
    assert same(iq(H), want_H)
    nopilot, final, demod = take(96), take(96), take(192)
    q = np.float32(1 / np.sqrt(2.0))
    assert same(final.real, np.where(nopilot.real > 0, q, -q)) and same(final.imag, np.where(nopilot.imag > 0, q, -q))      # AGC_Receiver :852
    a, b = final.real, final.imag
    c = np.where((a > 0) & (b > 0), 0, np.where((a < 0) & (b > 0), 0, 1))
    d = np.where((a > 0) & (b > 0), 0, np.where((a < 0) & (b > 0), 1, np.where((a < 0) & (b < 0), 0, 1)))
    assert same(demod.real.reshape(2, 48, 2), np.stack([c, d], axis=-1).reshape(2, 48, 2))               # QPSK_Demodulator :873
    assert pos[0] == raw.size
