"""Stand-alone 64-point transform stage (ofdm_fft64 / ofdm_ifft64, k_fft64): time and HBM rate in both arithmetic modes.
usage: python tools/fft_probe.py [millions of windows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
lib, h = o.lib, o.h
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 8_000_000
x = torch.randn((n, 64, 2), dtype=torch.float32, device=o.device)
y = torch.empty_like(x)
for name, fn in (("fft64", lib.ofdm_fft64), ("ifft64", lib.ofdm_ifft64)):
    for mode, mname in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        o._check(fn(h, x.data_ptr(), y.data_ptr(), n, mode)); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            o._check(fn(h, x.data_ptr(), y.data_ptr(), n, mode))
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print("%-7s %-5s %.3f ms per %d windows = %.0f GB/s (1024 B per window)" % (name, mname, ms, n, n * 1024 / ms / 1e6), flush=True)
