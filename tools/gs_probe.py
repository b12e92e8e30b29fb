import sys, os
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev = o.device; lib, h = o.lib, o.h
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 1_000_000
cnt = o.new_counters(1)
bits = torch.randint(-2**31, 2**31 - 1, (n * 6,), dtype=torch.int32, device=dev)
frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
g = torch.randn((n, 320), dtype=torch.float32, device=dev)
power = torch.empty((n,), dtype=torch.float32, device=dev)
for gs in (0, 1):
    o.set_option("general_stream", gs)
    for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, 2, mode)
        ms = t(lambda: lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 8.0, n, 2, mode, cnt.data_ptr(), None))
        print("general_stream %d %s: %.3f ms" % (gs, name, ms))
