"""Replay rate and time of the fused Monte-Carlo kernels (k_mc_quad / k_mc_philox) in EXACT mode: python tools/mc_replays.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
N = 1_000_000
SNRS = [float(s) for s in range(21)]
for layout in (0, 1):
    o.set_option("stream_layout", layout)
    for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        cnt = o.new_counters(21)
        o.mc_sweep_points(7, 0, N, 2, 0, SNRS, None, mode, cnt); torch.cuda.synchronize()
        o.replayed_frames(reset=True)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        cnt.zero_(); a.record()
        o.mc_sweep_points(7, 0, N, 2, 0, SNRS, None, mode, cnt)
        b.record(); torch.cuda.synchronize()
        c = o.read_counters(cnt)
        print("layout %d %-5s %.3f ms  replayed %d (%.3f %%)  bit errors at 0/10 dB %d %d" % (
            layout, name, a.elapsed_time(b), o.replayed_frames(), 100.0 * o.replayed_frames() / (N * 21), c[0].bit_errors, c[10].bit_errors), flush=True)
