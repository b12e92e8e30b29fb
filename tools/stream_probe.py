import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mode = pkg.MODE_FAST if (len(sys.argv) < 3 or sys.argv[2] == "fast") else pkg.MODE_EXACT
dev = o.device
bits = torch.randint(-2**31, 2**31-1, (n*6,), dtype=torch.int32, device=dev)
frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
cnt = o.new_counters(1)
lib, h = o.lib, o.h
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
ms_tx = t(lambda: lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, n, 2, mode))
ms_rx = t(lambda: lib.ofdm_rx_frames(h, frames.data_ptr(), bits.data_ptr(), n, 2, mode, cnt.data_ptr(), None))
print("frames %d mode %d: tx %.3f ms (%.0f GB/s)  rx %.3f ms (%.0f GB/s)" % (n, mode, ms_tx, n*2584/ms_tx/1e6, ms_rx, n*2072/ms_rx/1e6))
