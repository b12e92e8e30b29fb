"""BASELINE.json's full sizes (1,000,000 frames x 2 symbols, 21 SNR points) through size-independent properties:
noise-free round trip, chunk-sum invariance, agreement of the three routes to the same totals (fused sweep, staged
kernels, host-buffer sweep), monotone BER, exact == fast up to marginal decisions, and the 16 Mi-symbol streaming
round trip of configs[2]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, NSYM = 1_000_000, 2
SNRS = [float(s) for s in range(0, 21)]


def ints(c):
    return (c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames)


def test_noise_free_round_trip_full_size(ofdm, pkg):
    t = ofdm.torch
    bits = ofdm.random_bits(5, 0, N, NSYM)
    for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
        frames = ofdm.tx_frames(bits, NSYM, mode, with_power=False)
        cnt, d = ofdm.rx_frames(frames, bits, NSYM, mode, want=("bits", "frame_bit_errors"))
        assert cnt.bit_errors == 0 and cnt.frames == N and cnt.bits == N * 96 * NSYM and cnt.frames_in_error == 0
        assert t.equal(d["bits"].reshape(-1), bits.reshape(-1))            # encode -> decode is the identity on 192 M bits
        assert int(d["frame_bit_errors"].abs().sum()) == 0
        cnt2, _ = ofdm.rx_frames(frames, bits, NSYM, mode)                  # TMA-staged route
        assert ints(cnt2) == ints(cnt)
        evm = np.sqrt(cnt2.sum_err2 / cnt2.sum_ref2)
        assert evm < 2e-6                                                   # nothing but fp32 rounding
        # flipping one payload bit of one frame is seen as exactly one error (checks the comparison itself)
        b2 = bits.clone(); b2[3 * 2 * 777 + 1] ^= 1 << 7
        cnt3, _ = ofdm.rx_frames(frames, b2, NSYM, mode)
        assert cnt3.bit_errors == 1 and cnt3.frames_in_error == 1
        del frames, d


def test_sweep_routes_agree_full_size(ofdm, pkg):
    t = ofdm.torch
    gen = t.Generator(device=ofdm.device); gen.manual_seed(3)
    bits = ofdm.random_bits(9, 0, N, NSYM)
    g = t.randn((N, 320), dtype=t.float32, device=ofdm.device, generator=gen)
    whole = ofdm.sweep_inject_dev(bits, g, N, NSYM, SNRS, pkg.MODE_EXACT)
    ber = [c.bit_errors / c.bits for c in whole]
    assert all(a >= b for a, b in zip(ber, ber[1:])) and 0.25 < ber[0] < 0.27 and ber[20] == 0.0
    assert all(c.frames == N and c.bits == N * 192 for c in whole)
    # checksum of checksums: any split of the frame range gives the same integer totals
    cut = 333_337
    a = ofdm.sweep_inject_dev(bits[:cut * 6], g[:cut], cut, NSYM, SNRS, pkg.MODE_EXACT)
    b = ofdm.sweep_inject_dev(bits[cut * 6:], g[cut:], N - cut, NSYM, SNRS, pkg.MODE_EXACT)
    for w, x, y in zip(whole, a, b):
        assert ints(w) == tuple(p + q for p, q in zip(ints(x), ints(y)))
    # host-buffer route (chunked, copy/compute pipelined) == resident route
    host = ofdm.sweep_inject_host(bits.cpu().numpy(), g.cpu().numpy(), N, NSYM, SNRS, pkg.MODE_EXACT)
    for w, h in zip(whole, host):
        assert ints(w) == ints(h)
    # the speculating EXACT kernel (default) == the reference's arithmetic on every frame, 21 x 192 M rail decisions
    ofdm.set_option("exact_speculation", 0)
    try:
        all_exact = ofdm.sweep_inject_dev(bits, g, N, NSYM, SNRS, pkg.MODE_EXACT)
    finally:
        ofdm.set_option("exact_speculation", 1)
    for w, e in zip(whole, all_exact):
        assert ints(w) == ints(e)
        assert abs(w.sum_err2 - e.sum_err2) <= 1e-6 * e.sum_err2
    # fast mode differs from exact only by decisions within fp32 rounding of zero
    fast = ofdm.sweep_inject_dev(bits, g, N, NSYM, SNRS, pkg.MODE_FAST)
    for w, f in zip(whole, fast):
        assert abs(int(w.bit_errors) - int(f.bit_errors)) <= max(20, 2e-5 * w.bit_errors)
        assert abs(w.sum_err2 - f.sum_err2) <= 1e-4 * w.sum_err2


def test_streaming_16mi_symbols(ofdm, pkg):
    """configs[2]: 8,388,608 frames = 16 Mi data symbols resident in HBM, TX then RX, nothing lost."""
    n = 8_388_608
    bits = ofdm.random_bits(1, 0, n, NSYM)
    frames = ofdm.tx_frames(bits, NSYM, pkg.MODE_FAST, with_power=False)
    cnt, _ = ofdm.rx_frames(frames, bits, NSYM, pkg.MODE_FAST)
    assert cnt.frames == n and cnt.bits == n * 192 and cnt.bit_errors == 0


def test_cfg1_against_the_compiled_reference_100k(pkg, lib, ref):
    """configs[1] against the reference itself at 100,000 frames x 21 SNR points: draws from the reference's own
    rand() / Box-Muller stream, oracle/_ref's ref_chain_sweep on all host cores vs ofdm_sweep_inject_host in EXACT
    mode -- every integer total equal at every point, EVM sums within 1e-5.  (tools/full_parity.py runs the same at
    the full 1,000,000 frames; its output is kept under profiles/.)"""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import full_parity
    res = full_parity.run(100_000, len(os.sched_getaffinity(0)), verbose=False)
    bad = [p for p in res["points"] if not p["ints_equal"] or p["evm_rel_diff"] > 1e-5 or p["evm_lin_sum_rel_diff"] > 1e-5]
    assert res["pass"] and not bad, bad
    assert res["points"][0]["gpu"][1] == 100_000 * 192 and res["points"][0]["gpu"][4] == 100_000
