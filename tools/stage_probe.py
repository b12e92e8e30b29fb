"""The stand-alone stage entry points (include/ofdm_b200.h) one by one: time and algorithmic HBM rate.
usage: python tools/stage_probe.py [millions of frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 2_000_000
n_sym = 2
peak = 6554.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = o.device


def timed(fn, reps=5):
    out = fn(); out = fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def stage(name, fn, nbytes):
    out, ms = timed(fn)
    print("%-44s %8.3f ms  %6.0f GB/s  %.2f of HBM peak" % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak), flush=True)
    return out


packed = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * n_sym * 3,), dtype=torch.int32, device=dev)
S = n * n_sym
mod = stage("qpsk_modulate (a1)", lambda: o.qpsk_modulate(packed), S * (12 + 48 * 8))
grid = stage("map_subcarriers (a2)", lambda: o.map_subcarriers(mod), S * (48 * 8 + 64 * 8))
del mod
for mode, mn in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    t = stage("ifft64 %s (a3)" % mn, lambda: o.ifft64(grid, mode), S * 1024)
sym = stage("add_cp (a4)", lambda: o.add_cp(t), S * (512 + 640))
del grid, t, sym
torch.cuda.empty_cache()
frames, power = o.tx_frames(packed, n_sym, pkg.MODE_EXACT)
g = torch.randn((n, 320), dtype=torch.float32, device=dev)
for mode, mn in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    stage("frame_power %s (a7)" % mn, lambda: o.frame_power(frames, mode), n * 2564)
    ota = stage("awgn_inject %s (a7)" % mn, lambda: o.awgn_inject(frames, g, 8.0, n_sym, mode, power=power), n * (2560 * 2 + 1280 + 4))
del g
for mode, mn in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    H = stage("channel_estimate %s (a8)" % mn, lambda: o.channel_estimate(ota, mode), n * (1024 + 512))
bodies = stage("strip_cp (a9)", lambda: o.strip_cp(ota, n_sym), S * 1024)
del frames, ota
torch.cuda.empty_cache()
for mode, mn in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    F = stage("fft64 %s (a10)" % mn, lambda: o.fft64(bodies, mode), S * 1024).view(n, n_sym, 64, 2)
for mode, mn in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
    E = stage("equalize %s (a11)" % mn, lambda: o.equalize(F, H, mode), S * 1024 + n * 512)
del bodies, F, H
pts = stage("demap (a12)", lambda: o.demap(E), S * (512 + 384))
sl = stage("agc_slicer (a13)", lambda: o.agc_slicer(pts), S * 768)
stage("qpsk_demodulate (a14)", lambda: o.qpsk_demodulate(sl), S * (384 + 12))
