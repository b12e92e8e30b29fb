"""EVM guard trade-off: replays and distance of the EVM sums from the all-exact kernels for several guard widths.
    python tools/guard_probe.py [frames=1000000]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as e

pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
snr = [float(s) for s in range(0, 21, 2)]
gen = torch.Generator(device=o.device); gen.manual_seed(3)
bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * 6,), dtype=torch.int32, device=o.device, generator=gen)
g = torch.randn((n, 320), dtype=torch.float32, device=o.device, generator=gen)
o.set_option("exact_speculation", 0)
ref = o.sweep_inject_dev(bits, g, n, 2, snr, pkg.MODE_EXACT)
o.set_option("exact_speculation", 1)
for guard in (3280, 820, 410, 205, 102, 51, 13):
    o.set_option("evm_guard", guard)
    for fused in (1, 0):
        o.set_option("fused_sweep", fused)
        o.replayed_frames(reset=True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        got = o.sweep_inject_dev(bits, g, n, 2, snr, pkg.MODE_EXACT)
        e1.record(); torch.cuda.synchronize()
        rep = o.replayed_frames()
        same = all((a.bit_errors, a.rail_errors, a.frames_in_error) == (b.bit_errors, b.rail_errors, b.frames_in_error) for a, b in zip(got, ref))
        d2 = max(abs(a.sum_err2 - b.sum_err2) / b.sum_err2 for a, b in zip(got, ref))
        dl = max(abs(a.sum_evm_lin - b.sum_evm_lin) / b.sum_evm_lin for a, b in zip(got, ref))
        print("guard %4d  %s  replayed %8d of %d (%.3f %%)  ints equal %s  max rel diff sum_err2 %.2e  sum_evm_lin %.2e  sweep %.2f ms" %
              (guard, "all-SNR kernel " if fused else "kernel per point", rep, n * len(snr), 100.0 * rep / (n * len(snr)), same, d2, dl, e0.elapsed_time(e1)), flush=True)
