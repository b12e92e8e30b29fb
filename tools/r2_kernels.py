"""A few launches of each kernel of interest, for ncu captures and quick timings:
    python tools/r2_kernels.py <what> [reps]
what: sweep (k_sweep_lin<checked>) | sweep_fast | point (k_stream_rx2<checked,inject>) | rx_fast | rx_exact (k_stream_rx2<.,none>)
      | tx_fast | tx_exact (k_tx_frames2) | power | power_full | power_allref (k_frame_power_tiled) | mc_fast | mc_exact (k_mc_philox) | mp_fast | mp_exact | all (timings of all of them)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as e

pkg = e.load_pkg()
o = pkg.Ofdm(0)
dev, lib, h = o.device, o.lib, o.h
if os.environ.get("STREAM_LAYOUT"):          # 0 = one frame per lane group (k_stream_quad), 1 = one frame per warp (k_stream_rx2)
    o.set_option("stream_layout", int(os.environ["STREAM_LAYOUT"]))
if os.environ.get("STREAM_WARPS"):           # k_stream_quad: 6 (168 registers, 12 warps per SM) or 8 (128 registers, 16 warps per SM) warps per block
    o.set_option("stream_warps", int(os.environ["STREAM_WARPS"]))
what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SNRS = [float(s) for s in range(21)]
snr_arr = np.ascontiguousarray(SNRS, dtype=np.float32)
N = 1_000_000


def timed(fn, n=reps):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def setup(n, mode):
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * 6,), dtype=torch.int32, device=dev)
    g = torch.randn((n, 320), dtype=torch.float32, device=dev)
    frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
    power = torch.empty((n,), dtype=torch.float32, device=dev)
    o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, 2, mode))
    return bits, g, frames, power


def run(name):
    mode = pkg.MODE_FAST if name.endswith("fast") else pkg.MODE_EXACT
    if name in ("sweep", "sweep_fast"):
        bits, g, frames, power = setup(N, mode)
        cnt = o.new_counters(21)
        ms = timed(lambda: o._check(lib.ofdm_awgn_rx_inject_sweep(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(),
                                                                   snr_arr.ctypes.data, 21, N, 2, mode, cnt.data_ptr())))
        return ms, "%.3e frame-points/s" % (N * 21 / ms * 1e3)
    if name in ("point", "point_fast"):
        bits, g, frames, power = setup(N, mode)
        cnt = o.new_counters(1)
        ms = timed(lambda: o._check(lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), 6.0, N, 2, mode,
                                                             cnt.data_ptr(), None)))
        return ms, "%.0f GB/s" % (3100 * N / ms / 1e6)
    if name in ("rx_fast", "rx_exact", "tx_fast", "tx_exact"):
        n = 8_388_608
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * 6,), dtype=torch.int32, device=dev)
        frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
        cnt = o.new_counters(1)
        tx = lambda: o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, n, 2, mode))
        if name.startswith("tx"):
            ms = timed(tx)
            return ms, "%.0f GB/s" % (n * 2584 / ms / 1e6)
        tx()
        ms = timed(lambda: o._check(lib.ofdm_rx_frames(h, frames.data_ptr(), bits.data_ptr(), n, 2, mode, cnt.data_ptr(), None)))
        return ms, "%.0f GB/s" % (n * 2072 / ms / 1e6)
    if name in ("power", "power_full", "power_allref"):
        # the exact power chain: of transmitter frames (LTS prefix skipped: 1280 B read per frame) as the sweep runs it (timed as
        # transmitter + power minus transmitter alone), or of arbitrary frames (2560 B), or with every sample through hypot()
        bits, g, frames, power = setup(N, pkg.MODE_EXACT)
        if name == "power":
            a = timed(lambda: o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), N, 2, pkg.MODE_EXACT)))
            b = timed(lambda: o._check(lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, N, 2, pkg.MODE_EXACT)))
            return a - b, "%.0f GB/s (transmitter alone %.4f ms)" % (N * 1284 / (a - b) / 1e6, b)
        o.set_option("power_margin", 1 << 28 if name == "power_allref" else 16)
        ms = timed(lambda: o._check(lib.ofdm_frame_power(h, frames.data_ptr(), power.data_ptr(), N, 320, pkg.MODE_EXACT)))
        o.set_option("power_margin", 16)
        return ms, "%.0f GB/s" % (N * 2564 / ms / 1e6)
    if name in ("mc_fast", "mc_exact", "mp_fast", "mp_exact"):
        cnt = o.new_counters(21)
        taps = 8 if name.startswith("mp") else 0
        if taps:
            o.set_option("multipath_path", 2)
        ms = timed(lambda: o.mc_sweep_points(7, 0, N, 2, taps, SNRS, None, mode, cnt), max(1, reps // 2))
        return ms, "%.3e symbols/s" % (N * 2 * 21 / ms * 1e3)
    raise SystemExit("unknown kernel " + name)


names = ["power", "power_full", "power_allref", "sweep", "sweep_fast", "point", "point_fast", "rx_fast", "rx_exact", "tx_fast", "tx_exact", "mc_fast", "mc_exact", "mp_fast", "mp_exact"] if what == "all" else [what]
for nm in names:
    ms, rate = run(nm)
    print("%-12s %8.4f ms  %s" % (nm, ms, rate), flush=True)
    torch.cuda.empty_cache()
