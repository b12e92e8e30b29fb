"""What a replay costs in the all-SNR kernel: the sweep with every (frame, point) replayed against the normal sweep."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev, lib, h = o.device, o.lib, o.h
N = 200_000
SNRS = [float(s) for s in range(21)]
snr_arr = np.ascontiguousarray(SNRS, dtype=np.float32)
bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (N * 6,), dtype=torch.int32, device=dev)
g = torch.randn((N, 320), dtype=torch.float32, device=dev)
frames = torch.empty((N, 320, 2), dtype=torch.float32, device=dev)
power = torch.empty((N,), dtype=torch.float32, device=dev)
o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), N, 2, pkg.MODE_EXACT))
cnt = o.new_counters(21)
def run():
    o._check(lib.ofdm_awgn_rx_inject_sweep(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr_arr.ctypes.data, 21, N, 2, pkg.MODE_EXACT, cnt.data_ptr()))
def timed(n=5):
    run(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): run()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t0 = timed()
o.replayed_frames(reset=True); run(); rate = o.replayed_frames() / (N * 21)
o.set_option("force_replay", 1)
t1 = timed()
o.set_option("force_replay", 0)
print("N=%d x 21 points: normal %.4f ms (%.3f%% of the points replayed), every point replayed %.4f ms -> %.2f ns of kernel time per replay = %.1f speculated points; replays are %.1f%% of the normal kernel"
      % (N, t0, 100 * rate, t1, (t1 - t0) / (N * 21) * 1e6, (t1 - t0) / t0, 100 * rate * (t1 - t0) / t0))
