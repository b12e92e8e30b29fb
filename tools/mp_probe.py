"""cfg4 timing matrix: fused on-chip kernel vs staged path, fast / exact (checked) / exact (all-exact), with replay counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = 1_000_000; snr = [float(s) for s in range(21)]
cnt = o.new_counters(len(snr))
def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for staged in (0, 1):
    o.set_option("multipath_path", 1 if staged else 2)
    for name, mode, spec in (("fast", pkg.MODE_FAST, 1), ("exact checked", pkg.MODE_EXACT, 1), ("exact all", pkg.MODE_EXACT, 0)):
        o.set_option("exact_speculation", spec)
        o.replayed_frames(reset=True)
        ms = t(lambda: o.mc_sweep_multipath(11, 0, n, 2, 8, snr, mode, counters=cnt))
        print("%-7s %-14s %8.2f ms per 1M frames x 21   replays per sweep %d" % ("staged" if staged else "fused", name, ms, o.replayed_frames() // 3))
