"""CUDA path against the committed golden vectors (generated from the compiled reference by
tests/golden/make_golden.py): bit-exact in EXACT mode, stage by stage and for whole-chain totals; plus the
edge cases (degenerate frames, empty batches, bad arguments) and the C host driver."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits_and_noise

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_golden_stage_vectors_exact(ofdm, pkg, golden):
    g = golden
    lf, lt = ofdm.lts()
    assert same(lf, g["lts_freq"]) and same(lt, g["lts_time"])
    x = ofdm.to_dev(g["fft_in"])
    assert same(ofdm.fft64(x, pkg.MODE_EXACT).cpu().numpy(), g["fft_out"])
    assert same(ofdm.ifft64(x, pkg.MODE_EXACT).cpu().numpy(), g["ifft_out"])
    for n_sym in (2, 5):
        t = "n%d_" % n_sym
        bits = g[t + "bits"]
        packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
        mod = ofdm.qpsk_modulate(packed)
        assert same(mod.cpu().numpy().reshape(g[t + "mod"].shape), g[t + "mod"])
        grid = ofdm.map_subcarriers(mod)
        assert same(grid.cpu().numpy().reshape(g[t + "grid"].shape), g[t + "grid"])
        assert same(ofdm.ifft64(grid, pkg.MODE_EXACT).cpu().numpy().reshape(g[t + "sym_time"].shape), g[t + "sym_time"])
        frames, power = ofdm.tx_frames(packed, n_sym, pkg.MODE_EXACT)
        assert same(frames.cpu().numpy(), g[t + "tx"]) and same(power.cpu().numpy(), g[t + "power"])
        gd = ofdm.to_dev(g[t + "g"])
        for i, snr in enumerate(g[t + "snr"]):
            ota = ofdm.awgn_inject(frames, gd, float(snr), n_sym, pkg.MODE_EXACT, power=power)
            assert same(ota.cpu().numpy(), g[t + "ota_%d" % i])
            cnt, d = ofdm.rx_frames(ota, packed, n_sym, pkg.MODE_EXACT, want=("H", "eq", "sliced", "bits", "frame_bit_errors", "frame_evm_lin"))
            assert same(d["H"].cpu().numpy(), g[t + "rx_H_%d" % i])
            assert same(d["eq"].cpu().numpy(), g[t + "rx_eq_%d" % i])
            assert same(d["sliced"].cpu().numpy(), g[t + "rx_sliced_%d" % i])
            assert same(pkg.unpack_bits_host(d["bits"].cpu().numpy().view(np.uint32)), g[t + "rx_bits_%d" % i])
            assert same(d["frame_bit_errors"].cpu().numpy(), g[t + "rx_bit_errors_%d" % i])
            assert np.allclose(d["frame_evm_lin"].cpu().numpy(), g[t + "rx_evm_lin_%d" % i], rtol=1e-5)
            assert cnt.rail_errors == int(g[t + "rx_rail_errors_%d" % i].sum())


@pytest.mark.parametrize("mode", [0, 1])
def test_golden_chain_totals(ofdm, pkg, golden, mode):
    g = golden
    bits = pkg.pack_bits_host(g["chain_bits"])
    snrs = [float(s) for s in g["chain_snr"]]
    for res in (ofdm.sweep_inject_host(bits, g["chain_g"], 256, 2, snrs, mode),
                ofdm.sweep_inject_dev(ofdm.to_dev(bits.view(np.int32)), ofdm.to_dev(g["chain_g"]), 256, 2, snrs, mode)):
        for c, want in zip(res, g["chain_totals"]):
            if mode == pkg.MODE_EXACT:
                assert [c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames] == [int(v) for v in want[:5]]
            else:
                assert abs(int(c.bit_errors) - int(want[0])) <= 2 and c.bits == int(want[1])
            assert abs(c.sum_err2 / c.sum_ref2 - want[5] / want[6]) <= 2e-5 * want[5] / want[6]
            assert abs(c.sum_evm_lin - want[7]) <= 2e-5 * want[7]


def test_degenerate_frames_match_oracle(ofdm, pkg, port):
    """all-zero LTS (H = 0 -> libgcc's division recovery decides by the numerator's parts), all-zero frames,
    huge and tiny scalings: EXACT mode reproduces the reference's decisions, in the dump and in the sweep path."""
    bits, _ = bits_and_noise(9, 40, 2)
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    tx = port.tx_frames(bits, 2)
    z = tx.copy(); z[:, :160] = 0
    for frames in (z, np.zeros_like(tx), tx * np.float32(1e18), tx * np.float32(1e-18), tx):
        want = port.rx_frames(frames, bits, 2)
        cnt, d = ofdm.rx_frames(ofdm.to_dev(frames), packed, 2, pkg.MODE_EXACT, want=("bits", "frame_bit_errors"))
        assert same(pkg.unpack_bits_host(d["bits"].cpu().numpy().view(np.uint32)), want["bits"])
        assert same(d["frame_bit_errors"].cpu().numpy(), want["bit_errors"])
        cnt2, _ = ofdm.rx_frames(ofdm.to_dev(frames), packed, 2, pkg.MODE_EXACT)          # sweep path (no dump)
        assert cnt2.bit_errors == int(want["bit_errors"].sum()) and cnt2.rail_errors == int(want["rail_errors"].sum())


def test_empty_and_invalid_inputs(ofdm, pkg):
    t = ofdm.torch
    e32 = ofdm.empty((0,), t.int32)
    ef = ofdm.empty((0, 320, 2), t.float32)
    assert ofdm.pack_bits(ofdm.empty((0,), t.uint8)).numel() == 0
    assert ofdm.qpsk_modulate(e32).numel() == 0
    frames, power = ofdm.tx_frames(e32, 2, pkg.MODE_EXACT)
    assert frames.shape[0] == 0
    cnt, _ = ofdm.rx_frames(ef, e32, 2, pkg.MODE_EXACT)
    assert cnt.frames == 0 and cnt.bits == 0
    assert ofdm.sweep_inject_host(np.zeros((0, 6), np.uint32), np.zeros((0, 320), np.float32), 0, 2, [5.0], 0)[0].frames == 0
    assert ofdm.mc_sweep_philox(1, 0, 0, 2, [5.0], 0)[0].frames == 0
    lib, h = ofdm.lib, ofdm.h
    x = ofdm.zeros((4, 64, 2), t.float32)
    assert lib.ofdm_fft64(h, x.data_ptr(), x.data_ptr(), 4, 0) == 1            # in-place is refused
    assert lib.ofdm_fft64(h, x.data_ptr(), None, 4, 0) == 1
    assert lib.ofdm_fft64(h, x.data_ptr(), x.data_ptr(), 4, 7) == 1            # unknown mode
    assert lib.ofdm_tx_frames(h, e32.data_ptr(), x.data_ptr(), None, 1, 0, 0) == 1      # n_sym < 1
    assert lib.ofdm_tx_frames(h, None, x.data_ptr(), None, 1, 2, 0) == 1
    assert b"invalid argument" in lib.ofdm_last_error(h)
    assert lib.ofdm_mc_sweep_philox_dev(h, 1, 0, 10, 2, None, 3, 0, None) == 1


def test_c_host_driver_writes_reference_format(tmp_path):
    """host/ofdm_main.c: the default run (reference message, 35 points 6..40 dB) and a batched run."""
    exe = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "ofdm_sweep")
    if not os.path.exists(exe):
        import __graft_entry__ as entry
        entry.build()
    out = tmp_path / "data"
    out.mkdir()
    r = subprocess.run([exe, "--stage-chain", "--outdir", str(out), "--dump", str(tmp_path / "Code_Output")], capture_output=True, timeout=120)
    stdout = r.stdout.decode("latin-1")          # low-SNR points print garbled bytes, as the reference does
    assert r.returncode == 0, stdout + r.stderr.decode("latin-1")
    assert "Code Run Successful!" in stdout
    assert "Received Message: \nHey! I am Vivaswan" in stdout                   # decoded at the high-SNR points
    rows = {}
    for name in ("Output_SNR", "Output_EVM_AGC", "Output_EVM_AGC_DB", "Output_BER"):
        text = open(out / (name + ".txt")).read()
        assert text.endswith("\n") and text.count("\n") == 1
        rows[name] = [float(w) for w in text.split()]
        assert len(rows[name]) == 35
    assert rows["Output_SNR"] == [float(s) for s in range(6, 41)]
    assert rows["Output_BER"][-1] == 0.0 and rows["Output_EVM_AGC_DB"][-1] == float("-inf")
    ev = np.array(rows["Output_EVM_AGC"])
    assert abs(ev[-1] + 40) < 3 and abs(ev[14] + 20) < 3                        # EVM before the slicer tracks -SNR
    assert len(open(str(tmp_path / "Code_Output_real.txt")).read().split()) == 320
    assert len(open(str(tmp_path / "Code_Output_complex.txt")).read().split()) == 3 * 320
    # default run = the reference's whole over-the-air path (STS, RRC, x10, capture, detection, CFO): like OFDM.exe, most
    # points above ~9 dB decode the message; a late packet in the capture window can garble a point (as in the reference)
    r = subprocess.run([exe, "--outdir", str(out)], capture_output=True, timeout=120)
    stdout = r.stdout.decode("latin-1")
    assert r.returncode == 0, stdout + r.stderr.decode("latin-1")
    assert stdout.count("Received Message: \nHey! I am Vivaswan") >= 20
    ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
    evm = [float(w) for w in open(out / "Output_EVM_AGC.txt").read().split()]
    assert len(ber) == 35 and sum(b == 0.0 for b in ber[6:]) >= 22
    good = [e for e, b in zip(evm[14:], ber[14:]) if b == 0.0]
    assert all(-45 < e < -15 for e in good)                                     # EVM before the slicer ~ -SNR - 1 dB
    r = subprocess.run([exe, "--quiet", "--outdir", str(out), "--frames", "200000", "--snr-start", "0", "--snr-count", "11", "--snr-step", "2",
                        "--mode", "fast"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    ber = [float(w) for w in open(out / "Output_BER.txt").read().split()]
    assert len(ber) == 11 and abs(ber[0] - 0.261) < 0.003 and abs(ber[5] - 1.75e-3) < 2e-4 and ber[-1] == 0.0


def test_c_host_driver_multi_gpu_nccl(tmp_path):
    """--gpus N: frames sharded by global index + one NCCL all-reduce per buffer type; the result files must be
    byte-identical to the single-GPU run (integer totals are split-invariant, EXACT mode)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "ofdm_sweep")
    if not os.path.exists(exe):
        import __graft_entry__ as entry
        entry.build()
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("data%d" % gpus)
        out.mkdir()
        r = subprocess.run([exe, "--quiet", "--outdir", str(out), "--gpus", str(gpus), "--frames", "300001", "--snr-start", "0",
                            "--snr-count", "9", "--snr-step", "2"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append({n: open(out / n).read() for n in ("Output_BER.txt", "Output_EVM_AGC_DB.txt", "Output_SNR.txt")})
    assert outs[0] == outs[1]


def test_config0_matlab_output_bits_through_gpu_and_dump(ofdm, pkg, lib, golden, tmp_path):
    """SURVEY 8(c) config-0 check on the CUDA path: data/Matlab_Output.txt's 96 bits as a symbol payload, chain at high
    SNR, demodulated bits dumped in the reference's real-part format (what compare_double.py reads): max error 0."""
    b96 = golden["matlab_output_bits"]
    bits = np.concatenate([b96, b96[::-1]])[None, :]
    packed = ofdm.to_dev(pkg.pack_bits_host(bits).view(np.int32))
    frames, power = ofdm.tx_frames(packed, 2, pkg.MODE_EXACT)
    _, d = ofdm.awgn_rx_philox(frames, packed, 30.0, 1, 0, 0, 2, pkg.MODE_EXACT, power=power, want=("bits",))
    rx = pkg.unpack_bits_host(d["bits"].cpu().numpy().view(np.uint32))[0, :96]
    iq = np.stack([rx.astype(np.float32), np.zeros(96, np.float32)], axis=1)
    path = str(tmp_path / "Code_Output.txt")
    assert pkg.write_complex_array_to_file(lib, iq, path, 0) == 0
    code = np.array([float(w) for w in open(path).read().split()])
    assert code.size == 96 and np.max(np.abs(code - b96)) == 0.0
