"""One launch of a secondary kernel for ncu: `one_launch.py rxn <n_sym> <exact|fast>` `one_launch.py mp <exact|fast>` or `one_launch.py detect`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev = o.device; lib, h = o.lib, o.h
what = sys.argv[1]
mode = pkg.MODE_EXACT if sys.argv[-1] == "exact" else pkg.MODE_FAST
if what == "rxn":
    n_sym = int(sys.argv[2]); n = 4_000_000 // (2 + n_sym); flen = 160 + 80 * n_sym
    cnt = o.new_counters(1)
    bits = torch.randint(-2**31, 2**31 - 1, (n * 3 * n_sym,), dtype=torch.int32, device=dev)
    frames = torch.empty((n, flen, 2), dtype=torch.float32, device=dev)
    g = torch.randn((n, flen), dtype=torch.float32, device=dev)
    power = torch.empty((n,), dtype=torch.float32, device=dev)
    lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, n_sym, mode)
    for _ in range(3):
        lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 8.0, n, n_sym, mode, cnt.data_ptr(), None)
elif what == "detect":
    cap = torch.randn((32768, 3008, 2), dtype=torch.float32, device=dev)
    for _ in range(3):
        o.packet_detect(cap)
else:
    snr = [float(s) for s in range(21)]
    cnt = o.new_counters(21)
    o.set_option("multipath_path", 2)
    for _ in range(2):
        o.mc_sweep_multipath(11, 0, 1_000_000, 2, 8, snr, mode, counters=cnt)
torch.cuda.synchronize()
print("ok")
