import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
snrs = [float(s) for s in range(0, 21)]
for mode, name in ((1, "fast"), (0, "exact")):
    n = 2_000_000
    cnt = o.new_counters(len(snrs))
    o.mc_sweep_philox(1, 0, 100000, 2, snrs, mode, counters=cnt); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); o.mc_sweep_philox(1, 0, n, 2, snrs, mode, counters=cnt); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print("mc %s: %d frames x %d snr in %.2f ms -> %.3e symbols/s, %.3e frame-passes/s" % (name, n, len(snrs), ms, n*2*len(snrs)/ms*1e3, n*len(snrs)/ms*1e3))
c = o.read_counters(cnt)
print([round(x.bit_errors / x.bits, 6) for x in c])
