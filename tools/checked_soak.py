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
    for snr in snrs:
        res = []
        for spec in (1, 0):
            o.set_option("exact_speculation", spec)
            cnt.zero_(); o.replayed_frames(reset=True)
            o._check(lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2,
                                             pkg.MODE_EXACT, cnt.data_ptr(), None))
            res.append((cnt.cpu().numpy().reshape(-1).copy(), o.replayed_frames()))
        a, b = res[0][0], res[1][0]
        same = bool((a[:5] == b[:5]).all())
        decisions += n * 192; replays += res[0][1]; mismatches += 0 if same else 1
        e2a, e2b = a.view(np.float64)[5], b.view(np.float64)[5]
        print("seed %d snr %5.1f: bit errors %11d  replayed %7d (%.3f%%)  totals equal %s  sum_err2 rel diff %.1e"
              % (seed, snr, int(a[0]), res[0][1], 100.0 * res[0][1] / n, same, abs(e2a / e2b - 1)), flush=True)
print("decisions compared: %.3e  launches with different totals: %d  replayed frames: %d" % (decisions, mismatches, replays))
