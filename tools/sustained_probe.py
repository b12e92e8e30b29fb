"""Burst vs sustained: the per-SNR-point kernel and a plain device copy, timed over bursts of increasing length
(CUDA events around the whole burst).  Shows whether a long back-to-back sequence runs slower than a short one on this box.
    python tools/sustained_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as e

pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev, lib, h = o.device, o.lib, o.h
N = 1_000_000
bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (N * 6,), dtype=torch.int32, device=dev)
g = torch.randn((N, 320), dtype=torch.float32, device=dev)
frames = torch.empty((N, 320, 2), dtype=torch.float32, device=dev)
power = torch.empty((N,), dtype=torch.float32, device=dev)
o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), N, 2, pkg.MODE_EXACT))
cnt = o.new_counters(1)
a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)      # 1 GiB each


def burst(fn, reps):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


point = lambda: o._check(lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 6.0, N, 2, pkg.MODE_EXACT, cnt.data_ptr(), None))
copy = lambda: b.copy_(a)
for reps in (3, 10, 30, 100, 300, 1000):
    ms_p = burst(point, reps)
    ms_c = burst(copy, max(1, reps // 2))
    print("burst of %4d launches: point kernel %.4f ms = %.0f GB/s (algorithmic 3100 B/frame);  1 GiB copy %.4f ms = %.0f GB/s (read + write)" %
          (reps, ms_p, 3100 * N / ms_p / 1e6, ms_c, 2 * a.numel() * 2 / ms_c / 1e6), flush=True)
# the same with a per-launch event pair, as bench.py times it
evs = []
for i in range(210):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(); point(); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
ts = [x.elapsed_time(y) for x, y in evs]
print("210 launches with an event pair each: mean %.4f ms, first 10 mean %.4f, last 10 mean %.4f" % (sum(ts) / len(ts), sum(ts[:10]) / 10, sum(ts[-10:]) / 10))
