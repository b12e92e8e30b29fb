"""Time the per-SNR-point channel+receiver kernel (injected draws) in both arithmetic modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = o.device
lib, h = o.lib, o.h
bits = torch.randint(-2**31, 2**31 - 1, (n * 6,), dtype=torch.int32, device=dev)
frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
power = torch.empty((n,), dtype=torch.float32, device=dev)
g = torch.randn((n, 320), dtype=torch.float32, device=dev)
cnt = o.new_counters(1)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    rc = lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, 2, mode); assert rc == 0
    for snr in (0.0, 10.0, 20.0):
        ms = t(lambda: lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2, mode, cnt.data_ptr(), None))
        print("%s snr %4.1f: %.4f ms per %d frames  %.0f GB/s (3100 B/frame)" % (name, snr, ms, n, n * 3100 / ms / 1e6))
# the EXACT rows above ran with speculation (default); now the all-exact kernel, and the replay statistics
o.set_option("exact_speculation", 0)
for snr in (0.0, 10.0):
    ms = t(lambda: lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2, pkg.MODE_EXACT, cnt.data_ptr(), None))
    print("all-exact snr %4.1f: %.4f ms" % (snr, ms))
o.set_option("exact_speculation", 1)
import numpy as np
for snr in (0.0, 5.0, 10.0, 15.0, 20.0):
    res = []
    for spec in (1, 0):
        o.set_option("exact_speculation", spec)
        cnt.zero_(); o.replayed_frames(reset=True)
        lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2, pkg.MODE_EXACT, cnt.data_ptr(), None)
        rep = o.replayed_frames()
        res.append((cnt.cpu().numpy().copy().reshape(-1), rep))
    a, b = res[0][0].view(np.uint64), res[1][0].view(np.uint64)
    print("snr %4.1f: replayed %d of %d (%.3f%%)  counts equal: %s  bit_errors %d  sum_err2 rel diff %.2e" % (
        snr, res[0][1], n, 100.0 * res[0][1] / n, bool((a[:5] == b[:5]).all()), int(a[0]),
        abs(res[0][0].view(np.float64)[5] / res[1][0].view(np.float64)[5] - 1)))
