import sys, os
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
dev = o.device; lib, h = o.lib, o.h
n_sym = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = pkg.MODE_EXACT if (len(sys.argv) < 3 or sys.argv[2] == "exact") else pkg.MODE_FAST
n = 4_000_000 // (2 + n_sym)
flen = 160 + 80 * n_sym
cnt = o.new_counters(1)
bits = torch.randint(-2**31, 2**31 - 1, (n * 3 * n_sym,), dtype=torch.int32, device=dev)
frames = torch.empty((n, flen, 2), dtype=torch.float32, device=dev)
g = torch.randn((n, flen), dtype=torch.float32, device=dev)
power = torch.empty((n,), dtype=torch.float32, device=dev)
lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, n_sym, mode)
for _ in range(3):
    lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 8.0, n, n_sym, mode, cnt.data_ptr(), None)
torch.cuda.synchronize()
print("ok", n)
