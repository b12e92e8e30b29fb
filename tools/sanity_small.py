"""Calls every kernel once at small sizes (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
rng = np.random.default_rng(0)
for n_sym, n in ((2, 67), (3, 21), (1, 9)):
    bits = rng.integers(0, 2, (n, 96 * n_sym), dtype=np.uint8)
    g = rng.standard_normal((n, 160 + 80 * n_sym)).astype(np.float32)
    packed = o.pack_bits(o.to_dev(bits)); o.unpack_bits(packed)
    mod = o.qpsk_modulate(packed); grid = o.map_subcarriers(mod)
    for m in (0, 1):
        t = o.ifft64(grid, m); o.fft64(t, m); o.add_cp(t)
        frames, power = o.tx_frames(packed, n_sym, m)
        gd = o.to_dev(g)
        ota = o.awgn_inject(frames, gd, 7.0, n_sym, m, power=power)
        o.awgn_philox(frames, 7.0, 1, 2, 3, n_sym, m)
        o.rx_frames(ota, packed, n_sym, m, want=("H", "eq", "sliced", "bits", "frame_bit_errors", "frame_evm_lin"))
        o.rx_frames(ota, packed, n_sym, m)
        o.awgn_rx_inject(frames, gd, packed, 7.0, n_sym, m, power=power)
        o.awgn_rx_philox(frames, packed, 7.0, 1, 2, 3, n_sym, m)
        o.sweep_inject_host(pkg.pack_bits_host(bits), g, n, n_sym, [3.0, 9.0], m)
        o.mc_sweep_philox(5, 11, n, n_sym, [3.0, 9.0, 12.0], m)
        o.mc_sweep_multipath(5, 0, n, n_sym, 5, [10.0], m)
        sh = o.rrc_tx(frames); o.rrc_rx(sh, 20, frames.shape[1]); o.awgn_inject_len(sh, o.to_dev(rng.standard_normal((n, sh.shape[1])).astype(np.float32)), 9.0, m)
    full = o.prepend_sts(frames)
    rep = o.gather(o.rrc_tx(full), 0, 4 * (2 * full.shape[1] + 20))
    cap = o.gather(rep, o.to_dev(rng.integers(0, 500, n).astype(np.int32)), 2 * full.shape[1] + 600)
    corr = o.packet_detect(cap); idx = o.packet_select(corr)
    rx = o.rrc_rx_idx(cap, idx, full.shape[1]); c1, _ = o.cfo(rx, False); o.cfo(c1, True)
    o.random_bits(1, 2, n, n_sym); o.frame_power(frames, 0); o.frame_power(frames, 1)
torch.cuda.synchronize()
print("sanity_small ok, launches", o.launch_count)
