"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.
EXACT mode: bit-exact IQ / decisions / counts.  FAST mode: 1e-5 relative (to the array's peak
magnitude) on IQ, 1e-5 relative on EVM; decisions may differ only where a rail sits within
rounding distance of zero."""
import os

import numpy as np
import pytest

from conftest import bits_and_noise

pytestmark = pytest.mark.gpu

REL = 1e-5   # north_star tolerance for fp32 IQ / EVM


def same(a, b):
    """numerically identical (+0 == -0, NaN == NaN)"""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def close_rel(a, b, rel=REL):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.max(np.abs(b)) if b.size else 1.0
    return np.max(np.abs(a - b)) <= rel * scale if b.size else True


@pytest.mark.parametrize("n_sym", [1, 2, 3, 7])
def test_stage_chain_tx(ofdm, pkg, port, n_sym):
    bits, _ = bits_and_noise(11 + n_sym, 37, n_sym)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    # layout helpers round-trip
    packed2 = ofdm.pack_bits(ofdm.to_dev(bits))
    assert same(packed2.cpu().numpy(), packed.cpu().numpy().reshape(-1))
    assert same(ofdm.unpack_bits(packed).cpu().numpy().reshape(bits.shape), bits)
    # a1 QPSK map, a2 subcarrier map
    mod = ofdm.qpsk_modulate(packed)
    want_mod = port.qpsk_mod(bits)
    assert same(mod.cpu().numpy(), want_mod)
    grid = ofdm.map_subcarriers(mod)
    want_grid = port.map_grid(want_mod)
    assert same(grid.cpu().numpy(), want_grid)
    # a3 ifft (exact: bit-identical incl. the 32-sample rotation; fast: 1e-5)
    want_t = port.ifft64(want_grid)
    t_exact = ofdm.ifft64(grid, pkg.MODE_EXACT)
    assert same(t_exact.cpu().numpy(), want_t)
    assert close_rel(ofdm.ifft64(grid, pkg.MODE_FAST).cpu().numpy(), want_t)
    # a4 CP
    cp = ofdm.add_cp(t_exact).cpu().numpy()
    assert same(cp[:, 16:], want_t) and same(cp[:, :16], want_t[:, 48:])
    # a5 LTS
    lf, lt = ofdm.lts()
    assert same(lf, port.lts_freq()) and same(lt, port.lts_time())
    # a1..a6 fused
    want_frames = port.tx_frames(bits, n_sym)
    frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
    assert same(frames.cpu().numpy(), want_frames)
    want_p = np.array([port.frame_power(f) for f in want_frames], np.float32)
    assert same(power.cpu().numpy(), want_p)
    ff, pf = ofdm.tx_frames(packed, n_sym, pkg.MODE_FAST)
    assert close_rel(ff.cpu().numpy(), want_frames)
    assert np.allclose(pf.cpu().numpy(), want_p, rtol=REL)


def test_fft_exact_random(ofdm, pkg, port):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((1000, 64, 2)) * np.exp(rng.uniform(-3, 3, (1000, 1, 1)))).astype(np.float32)
    xd = ofdm.to_dev(x)
    assert same(ofdm.fft64(xd, pkg.MODE_EXACT).cpu().numpy(), port.fft64(x))
    assert same(ofdm.ifft64(xd, pkg.MODE_EXACT).cpu().numpy(), port.ifft64(x))
    for i in range(0, 1000, 100):
        assert close_rel(ofdm.fft64(xd[i:i + 1], pkg.MODE_FAST).cpu().numpy(), port.fft64(x[i:i + 1]))
        assert close_rel(ofdm.ifft64(xd[i:i + 1], pkg.MODE_FAST).cpu().numpy(), port.ifft64(x[i:i + 1]))


@pytest.mark.parametrize("n_sym,snr", [(2, 0.0), (2, 6.0), (2, 12.0), (2, 30.0), (1, 4.0), (5, 7.0), (9, 9.0)])
def test_channel_and_receiver_exact(ofdm, pkg, port, n_sym, snr):
    n_frames = 300
    bits, g = bits_and_noise(100 + n_sym, n_frames, n_sym)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    gd = ofdm.to_dev(g)
    frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
    want_tx = port.tx_frames(bits, n_sym)
    # a7 AWGN with injected draws (power given and power computed internally)
    want_ota = port.awgn_inject(want_tx, g, snr)
    ota = ofdm.awgn_inject(frames, gd, snr, n_sym, pkg.MODE_EXACT, power=power)
    assert same(ota.cpu().numpy(), want_ota)
    ota2 = ofdm.awgn_inject(frames, gd, snr, n_sym, pkg.MODE_EXACT)
    assert same(ota2.cpu().numpy(), want_ota)
    # a8..a16 receiver on the OTA buffer
    want = port.rx_frames(want_ota, bits, n_sym)
    cnt, d = ofdm.rx_frames(ota, packed, n_sym, pkg.MODE_EXACT,
                            want=("H", "eq", "sliced", "bits", "frame_bit_errors", "frame_evm_lin"))
    assert same(d["H"].cpu().numpy(), want["H"])
    assert same(d["eq"].cpu().numpy(), want["eq"])
    assert same(d["sliced"].cpu().numpy(), want["sliced"])
    rx_bits = pkg.unpack_bits_host(d["bits"].cpu().numpy().view(np.uint32))
    assert same(rx_bits, want["bits"])
    assert same(d["frame_bit_errors"].cpu().numpy(), want["bit_errors"])
    assert np.allclose(d["frame_evm_lin"].cpu().numpy(), want["evm_lin"], rtol=REL)
    assert cnt.bit_errors == int(want["bit_errors"].sum())
    assert cnt.rail_errors == int(want["rail_errors"].sum())
    assert cnt.frames_in_error == int((want["bit_errors"] > 0).sum())
    assert cnt.frames == n_frames and cnt.bits == n_frames * 96 * n_sym
    # fused channel + receiver gives the same totals and per-frame results
    cnt2, d2 = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_EXACT, power=power,
                                   want=("frame_bit_errors", "eq"))
    assert same(d2["frame_bit_errors"].cpu().numpy(), want["bit_errors"])
    assert same(d2["eq"].cpu().numpy(), want["eq"])
    assert (cnt2.bit_errors, cnt2.rail_errors, cnt2.frames_in_error) == (cnt.bit_errors, cnt.rail_errors, cnt.frames_in_error)
    acc = port.chain(bits, g, n_sym, snr)
    assert cnt2.bit_errors == acc.bit_errors and cnt2.rail_errors == acc.rail_errors
    evm_gpu = np.sqrt(cnt2.sum_err2 / cnt2.sum_ref2); evm_cpu = np.sqrt(acc.sum_err2 / acc.sum_ref2)
    assert abs(evm_gpu - evm_cpu) <= REL * evm_cpu
    # Res[3]
    res = ofdm.finalize(cnt2)
    assert abs(res[2] - acc.bit_errors / acc.bits) < 1e-7
    assert abs(res[0] - 20 * np.log10(evm_cpu)) < 1e-3


@pytest.mark.parametrize("n_sym,snr", [(2, 3.0), (2, 10.0), (2, 25.0), (4, 8.0)])
def test_receiver_fast(ofdm, pkg, port, n_sym, snr):
    n_frames = 400
    bits, g = bits_and_noise(200 + n_sym, n_frames, n_sym)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    gd = ofdm.to_dev(g)
    want_tx = port.tx_frames(bits, n_sym)
    want_ota = port.awgn_inject(want_tx, g, snr)
    want = port.rx_frames(want_ota, bits, n_sym)
    frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_FAST)
    ota = ofdm.awgn_inject(frames, gd, snr, n_sym, pkg.MODE_FAST, power=power)
    assert close_rel(ota.cpu().numpy(), want_ota)
    cnt, d = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_FAST, power=power,
                                 want=("H", "eq", "frame_bit_errors", "frame_evm_lin"))
    assert close_rel(d["H"].cpu().numpy(), want["H"])
    # Per-bin dumps come from the generic kernel, plain fp32 end to end.  Equalised points relative to each frame's peak
    # magnitude: a bin whose channel estimate is nearly zero amplifies the fp32 transform's error in H (the measured worst
    # figures go to $OFDM_TEST_LOG when set; DESIGN.md section 4 quotes them).
    eq, weq = d["eq"].cpu().numpy().astype(np.float64), want["eq"].astype(np.float64)
    scale = np.abs(weq).reshape(n_frames, -1).max(axis=1)[:, None, None]
    worst_eq = np.max(np.abs(eq - weq) / scale)
    worst_evm = np.max(np.abs(d["frame_evm_lin"].cpu().numpy() / want["evm_lin"] - 1))
    if os.environ.get("OFDM_TEST_LOG"):
        with open(os.environ["OFDM_TEST_LOG"], "a") as f:
            f.write("test_receiver_fast n_sym=%d snr=%.1f worst |eq - ref| / frame peak %.3e, worst per-frame EVM rel %.3e\n" % (n_sym, snr, worst_eq, worst_evm))
    assert worst_eq <= REL and worst_evm <= REL, (worst_eq, worst_evm)        # measured on B200: 4.4e-6 / 3.3e-6 at 3 dB, < 1e-6 from 8 dB up
    diff = np.abs(d["frame_bit_errors"].cpu().numpy() - want["bit_errors"])
    assert diff.sum() <= 2          # decisions differ only for rails within fp32 rounding of zero
    # The EVM the path reports (batch totals, streaming kernels) meets the 1e-5 of the north star in FAST mode too: those
    # kernels replay the few frames with a tiny |H| bin in the reference's arithmetic (the EVM guard of ofdm_chain.cuh).
    evm_cpu = np.sqrt(np.sum(want["evm_lin"].astype(np.float64) ** 2) / n_frames)
    tot, _ = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_FAST, power=power)
    evm_gpu = np.sqrt(tot.sum_err2 / tot.sum_ref2)
    assert abs(evm_gpu - evm_cpu) <= REL * evm_cpu, (evm_gpu, evm_cpu)
    assert abs(tot.sum_evm_lin - float(np.sum(want["evm_lin"].astype(np.float64)))) <= REL * float(np.sum(want["evm_lin"].astype(np.float64)))
    assert abs(int(tot.bit_errors) - int(want["bit_errors"].sum())) <= 2
    sw = ofdm.sweep_inject_dev(packed, gd, n_frames, n_sym, [snr], pkg.MODE_FAST)[0]           # the all-SNR kernel (n_sym = 2) / per-point route
    assert abs(np.sqrt(sw.sum_err2 / sw.sum_ref2) - evm_cpu) <= REL * evm_cpu


@pytest.mark.parametrize("mode", [0, 1])
def test_streaming_kernel_equals_generic_kernel(ofdm, pkg, port, mode):
    """n_sym == 2 sweeps run the TMA-staged kernel (k_stream_rx2); it must give the generic kernel's totals
    (bit-identical integers in both modes: same arithmetic, different staging) for all three noise sources,
    including frame counts that do not fill the last chunk / stage ring."""
    for n_frames in (1, 2, 3, 777, 5000):
        bits, g = bits_and_noise(300 + n_frames, n_frames, 2)
        packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
        gd = ofdm.to_dev(g)
        frames, power = ofdm.tx_frames(packed, 2, mode)
        ota = ofdm.awgn_inject(frames, gd, 6.0, 2, mode, power=power)
        res = {}
        for generic in (1, 0):
            ofdm.set_option("force_generic_rx", generic)
            a, _ = ofdm.awgn_rx_inject(frames, gd, packed, 6.0, 2, mode, power=power)
            b, _ = ofdm.awgn_rx_philox(frames, packed, 6.0, 11, 3, 1000, 2, mode, power=power)
            c, _ = ofdm.rx_frames(ota, packed, 2, mode)
            res[generic] = (a, b, c)
        ofdm.set_option("force_generic_rx", 0)
        for x, y in zip(res[0], res[1]):
            assert (x.bit_errors, x.rail_errors, x.frames_in_error, x.frames, x.bits) == \
                   (y.bit_errors, y.rail_errors, y.frames_in_error, y.frames, y.bits)
            assert abs(x.sum_err2 - y.sum_err2) <= 1e-5 * y.sum_err2 and abs(x.sum_evm_lin - y.sum_evm_lin) <= 1e-4 * y.sum_evm_lin
        # inject == (awgn then rx) on the same draws
        assert res[0][0].bit_errors == res[0][2].bit_errors
        if mode == pkg.MODE_EXACT:
            assert res[0][0].bit_errors == port.chain(bits, g, 2, 6.0).bit_errors


@pytest.mark.parametrize("n_sym", [6, 37])
def test_long_frames_generic_receiver(ofdm, pkg, port, n_sym):
    """many symbols per frame: several passes of four windows, the last one partly empty (generic kernel)."""
    n_frames = 40
    bits, g = bits_and_noise(900 + n_sym, n_frames, n_sym)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    gd = ofdm.to_dev(g)
    frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
    assert same(frames.cpu().numpy(), port.tx_frames(bits, n_sym))
    for snr in (4.0, 11.0):
        acc, fe, fv = port.chain(bits, g, n_sym, snr, per_frame=True)
        cnt, d = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_EXACT, power=power, want=("frame_bit_errors", "frame_evm_lin"))
        assert same(d["frame_bit_errors"].cpu().numpy(), fe)
        assert np.allclose(d["frame_evm_lin"].cpu().numpy(), fv, rtol=REL)
        cnt2, _ = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_EXACT, power=power)      # no-dump path
        assert (cnt2.bit_errors, cnt2.rail_errors, cnt2.frames_in_error) == (acc.bit_errors, acc.rail_errors, acc.frames_in_error)
        cnt3, _ = ofdm.awgn_rx_inject(frames, gd, packed, snr, n_sym, pkg.MODE_FAST, power=power)
        assert abs(int(cnt3.bit_errors) - int(acc.bit_errors)) <= 3
