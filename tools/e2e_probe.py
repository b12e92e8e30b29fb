"""Where does the host sweep's time go?  ofdm_sweep_inject_host with 21 / 1 SNR points, exact / fast."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = 1_000_000
bits_h = torch.randint(-2**31, 2**31 - 1, (n, 6), dtype=torch.int32).pin_memory()
g_h = torch.randn((n, 320), dtype=torch.float32).pin_memory()
def run(snrs, mode, reps=5):
    o.sweep_inject_host(bits_h, g_h, n, 2, snrs, mode); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        o.sweep_inject_host(bits_h, g_h, n, 2, snrs, mode)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
full = [float(s) for s in range(21)]
for name, snrs, mode in (("exact 21 pts", full, pkg.MODE_EXACT), ("fast 21 pts", full, pkg.MODE_FAST), ("exact 1 pt", [10.0], pkg.MODE_EXACT),
                         ("fast 1 pt", [10.0], pkg.MODE_FAST), ("exact 10 pts", full[:10], pkg.MODE_EXACT)):
    print("%-14s %.3f ms per sweep" % (name, run(snrs, mode)))
