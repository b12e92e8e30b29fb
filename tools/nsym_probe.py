"""Receiver throughput per data symbol for other frame shapes (the generic kernel) against the default two-symbol frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev = o.device; lib, h = o.lib, o.h
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
cnt = o.new_counters(1)
for n_sym, generic in ((2, 0), (2, 1), (1, 0), (4, 0), (8, 0), (16, 0)):
    n = 4_000_000 // (2 + n_sym)
    flen = 160 + 80 * n_sym
    bits = torch.randint(-2**31, 2**31 - 1, (n * 3 * n_sym,), dtype=torch.int32, device=dev)
    frames = torch.empty((n, flen, 2), dtype=torch.float32, device=dev)
    g = torch.randn((n, flen), dtype=torch.float32, device=dev)
    power = torch.empty((n,), dtype=torch.float32, device=dev)
    o.set_option("force_generic_rx", generic)
    for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, n_sym, mode)
        ms = t(lambda: lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 8.0, n, n_sym, mode, cnt.data_ptr(), None))
        print("n_sym %2d %s %-5s: %.3f ms for %d frames = %.2e data symbols/s, %.2e windows/s" % (
            n_sym, "generic" if (generic or n_sym != 2) else "stream ", name, ms, n, n * n_sym / ms * 1e3, n * (2 + n_sym) / ms * 1e3))
    del bits, frames, g, power
