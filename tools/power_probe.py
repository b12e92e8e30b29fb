"""Is the sweep kernel power-limited when launched back to back?  Times isolated launches (GPU idle before each)
against a long back-to-back train while polling NVML clocks / power.  Analysis tool."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
import __graft_entry__ as e
pkg = e.load_pkg(); o = pkg.Ofdm(0)
n = 1_000_000
dev = o.device; lib, h = o.lib, o.h
bits = torch.randint(-2**31, 2**31 - 1, (n * 6,), dtype=torch.int32, device=dev)
frames = torch.empty((n, 320, 2), dtype=torch.float32, device=dev)
power = torch.empty((n,), dtype=torch.float32, device=dev)
g = torch.randn((n, 320), dtype=torch.float32, device=dev)
cnt = o.new_counters(1)
lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), power.data_ptr(), n, 2, pkg.MODE_EXACT)
pynvml.nvmlInit(); nh = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(nh, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(nh, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(nh) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(nh)))
        time.sleep(0.002)
th = threading.Thread(target=poll, daemon=True); th.start()
def launch(mode=pkg.MODE_EXACT, snr=10.0):
    lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), g.data_ptr(), power.data_ptr(), bits.data_ptr(), snr, n, 2, mode, cnt.data_ptr(), None)
def train(k, mode=pkg.MODE_EXACT):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    evs[0].record()
    for i in range(k):
        launch(mode); evs[i + 1].record()
    torch.cuda.synchronize()
    return np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(k)])
launch(); torch.cuda.synchronize()
for name, mode in (("checked", pkg.MODE_EXACT), ("fast", pkg.MODE_FAST)):
    iso = []
    for _ in range(5):
        time.sleep(0.4); iso.append(train(1, mode)[0])
    print(name, "isolated launches (0.4 s idle before each): ms", np.round(iso, 4))
    time.sleep(0.5)
    t0 = time.perf_counter(); tr = train(400, mode); t1 = time.perf_counter()
    print(name, "back-to-back x400: first 5", np.round(tr[:5], 4), " 20-25", np.round(tr[20:25], 4), " last 5", np.round(tr[-5:], 4), " mean %.4f" % tr.mean())
    s = [x for x in samples if t0 <= x[0] <= t1]
    if s:
        sm = np.array([x[1] for x in s]); mem = np.array([x[2] for x in s]); pw = np.array([x[3] for x in s])
        print("   during the train: %d samples, SM MHz min/median/max %d/%d/%d, MEM MHz min/max %d/%d, power W median/max %.0f/%.0f, reasons OR 0x%x"
              % (len(s), sm.min(), np.median(sm), sm.max(), mem.min(), mem.max(), np.median(pw), pw.max(), int(np.bitwise_or.reduce([x[4] for x in s]))))
stop = True
