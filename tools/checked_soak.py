"""Soak test of the speculating EXACT kernel: totals of the checked kernel against the all-exact kernel over many
frames, SNR points and seeds (every integer total must be identical).  usage: checked_soak.py [millions of frames per seed] [seeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 2_000_000
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
snrs = [-6.0, -3.0] + [float(s) for s in range(0, 21, 2)] + [30.0]
dev = o.device; lib, h = o.lib, o.h
frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
power = torch.empty((n,), dtype=torch.float32, device=dev)
cnt = o.new_counters(1)
decisions = mismatches = replays = 0
for seed in range(seeds):
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + seed)
    bits = torch.randint(-2**31, 2**31 - 1, (n * 6,), dtype=torch.int32, device=dev, generator=gen)
    g = torch.randn((n, 320), dtype=torch.float32, device=dev, generator=gen)
    lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, 2, pkg.MODE_EXACT)
    if seed % 3 == 1:
        frames *= 2.0 ** (seed + 3)            # exact rescaling of the waveform; the power follows
        o._check(lib.ofdm_frame_power(h, frames.data_ptr(), power.data_ptr(), n, 320, pkg.MODE_EXACT))
    exact_by_snr = []
    for snr in snrs:
        res = []
        for spec in (1, 0):
            o.set_option("exact_speculation", spec)
            cnt.zero_(); o.replayed_frames(reset=True)
            o._check(lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2,
                                             pkg.MODE_EXACT, cnt.data_ptr(), None))
            res.append((cnt.cpu().numpy().reshape(-1).copy(), o.replayed_frames()))
        a, b = res[0][0], res[1][0]
        exact_by_snr.append(b)
        same = bool((a[:5] == b[:5]).all())
        decisions += n * 192; replays += res[0][1]; mismatches += 0 if same else 1
        e2a, e2b = a.view(np.float64)[5], b.view(np.float64)[5]
        print("seed %d snr %5.1f: bit errors %11d  replayed %7d (%.3f%%)  totals equal %s  sum_err2 rel diff %.1e"
              % (seed, snr, int(a[0]), res[0][1], 100.0 * res[0][1] / n, same, abs(e2a / e2b - 1)), flush=True)
    # the all-SNR kernel (k_sweep_lin: one transform per window, a multiply-add per SNR point) against the same all-exact totals
    o.set_option("exact_speculation", 1)
    sw = o.new_counters(len(snrs)); o.replayed_frames(reset=True)
    snr_arr = np.ascontiguousarray(snrs, dtype=np.float32)
    o._check(lib.ofdm_awgn_rx_inject_sweep(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr_arr.ctypes.data, len(snrs), n, 2,
                                           pkg.MODE_EXACT, sw.data_ptr()))
    swn = sw.cpu().numpy(); rp = o.replayed_frames()
    same = all(bool((swn[i][:5] == exact_by_snr[i][:5]).all()) for i in range(len(snrs)))
    worst = max(abs(swn[i].view(np.float64)[5] / exact_by_snr[i].view(np.float64)[5] - 1) for i in range(len(snrs)))
    decisions += n * 192 * len(snrs); replays += rp; mismatches += 0 if same else 1
    print("seed %d all-SNR kernel: %d points replayed (%.3f%%)  totals equal at all %d SNR points %s  worst sum_err2 rel diff %.1e"
          % (seed, rp, 100.0 * rp / (n * len(snrs)), len(snrs), same, worst), flush=True)
print("decisions compared: %.3e  launches with different totals: %d  replayed frames: %d" % (decisions, mismatches, replays))
# the fused Monte-Carlo kernel, speculating vs all-exact, on the same Philox streams
for seed in range(seeds):
    res = []
    for spec in (1, 0):
        o.set_option("exact_speculation", spec); o.replayed_frames(reset=True)
        c = o.mc_sweep_philox(500 + seed, 0, n, 2, [float(s) for s in range(0, 21)], pkg.MODE_EXACT)
        res.append(([(x.bit_errors, x.bits, x.frames_in_error, x.rail_errors, x.frames) for x in c], [x.sum_err2 for x in c], o.replayed_frames()))
    print("Monte-Carlo seed %d: %d frames x 21 SNR points, replayed %d (%.2f%%), totals equal %s, worst sum_err2 rel diff %.1e"
          % (seed, n, res[0][2], 100.0 * res[0][2] / (n * 21), res[0][0] == res[1][0], max(abs(a / b - 1) for a, b in zip(res[0][1], res[1][1]))), flush=True)

# the same comparison on fading channels (configs[4]: 8 random taps per frame, on-chip Philox noise, HBM-staged path) and on
# other frame shapes (multi-pass streaming receiver)
snr21 = [float(s) for s in range(0, 21)]
for seed in range(seeds):
    res = []
    for spec in (1, 0):
        o.set_option("exact_speculation", spec); o.set_option("multipath_path", 1)
        o.replayed_frames(reset=True)
        c = o.mc_sweep_multipath(100 + seed, 0, n, 2, 8, snr21, pkg.MODE_EXACT)
        res.append(([(x.bit_errors, x.bits, x.frames_in_error, x.rail_errors, x.frames) for x in c], [x.sum_err2 for x in c], o.replayed_frames()))
    same = res[0][0] == res[1][0]
    worst = max(abs(a / b - 1) for a, b in zip(res[0][1], res[1][1]))
    print("multipath seed %d: %d frames x 21 SNR points, replayed %d (%.2f%%), totals equal %s, worst sum_err2 rel diff %.1e"
          % (seed, n, res[0][2], 100.0 * res[0][2] / (n * 21), same, worst), flush=True)
o.set_option("multipath_path", 0)
for n_sym in (1, 3, 5, 12):
    nf = n // (2 + n_sym) * 4
    gen = torch.Generator(device=dev); gen.manual_seed(77 + n_sym)
    b = torch.randint(-2**31, 2**31 - 1, (nf * 3 * n_sym,), dtype=torch.int32, device=dev, generator=gen)
    gg = torch.randn((nf, 160 + 80 * n_sym), dtype=torch.float32, device=dev, generator=gen)
    fr = torch.empty((nf, 160 + 80 * n_sym, 2), dtype=torch.float32, device=dev)
    pw = torch.empty((nf,), dtype=torch.float32, device=dev)
    lib.ofdm_tx_frames(h, b.data_ptr(), fr.data_ptr(), pw.data_ptr(), nf, n_sym, pkg.MODE_EXACT)
    for snr in (0.0, 4.0, 8.0):
        res = []
        for spec in (1, 0):
            o.set_option("exact_speculation", spec)
            cnt.zero_(); o.replayed_frames(reset=True)
            o._check(lib.ofdm_awgn_rx_inject(h, fr.data_ptr(), gg.data_ptr(), pw.data_ptr(), b.data_ptr(), snr, nf, n_sym, pkg.MODE_EXACT, cnt.data_ptr(), None))
            res.append((cnt.cpu().numpy().reshape(-1).copy(), o.replayed_frames()))
        a, bb = res[0][0], res[1][0]
        print("n_sym %2d snr %4.1f: %d frames, replayed %d, totals equal %s, sum_err2 rel diff %.1e"
              % (n_sym, snr, nf, res[0][1], bool((a[:5] == bb[:5]).all()), abs(a.view(np.float64)[5] / bb.view(np.float64)[5] - 1)), flush=True)
    del b, gg, fr, pw
o.set_option("exact_speculation", 1)
