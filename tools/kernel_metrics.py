"""profiles/r2_kernel_metrics.json from the per-kernel ncu summaries (tools/ncu_summary.py's .json next to each .txt):
    python tools/kernel_metrics.py <dir with <prefix>ncu_<what>.json> <prefix> <out.json>
Units per launch are those of tools/r2_kernels.py (1 M frames x 21 SNR points; 8 Mi frames for the cfg2 kernels)."""
import json
import os
import sys

src, prefix, out = sys.argv[1], sys.argv[2], sys.argv[3]
N, NS, N2 = 1_000_000, 21, 8_388_608
KERNELS = {   # what -> (key, units per launch, unit name, committed summary)
    "sweep": ("k_sweep_lin_checked", N * NS, "frame x SNR point"),
    "sweep_fast": ("k_sweep_lin_fast", N * NS, "frame x SNR point"),
    "point": ("k_stream_quad_checked_inject", N, "frame"),
    "point_fast": ("k_stream_quad_fast_inject", N, "frame"),
    "rx_fast": ("k_stream_quad_fast_none", N2, "frame"),
    "rx_exact": ("k_stream_quad_checked_none", N2, "frame"),
    "tx_fast": ("k_tx_frames2_fast", N2, "frame"),
    "tx_exact": ("k_tx_frames2_exact", N2, "frame"),
    "mc_fast": ("k_mc_quad_fast", N * NS, "frame x SNR point"),
    "mc_exact": ("k_mc_quad_checked", N * NS, "frame x SNR point"),
    "mp_fast": ("k_mc_quad_fast_multipath", N * NS, "frame x SNR point"),
    "power_full": ("k_frame_power_tiled", N, "frame (320 samples)"),
}


def num(s):
    return float(str(s).split()[0])


res = {}
for what, (key, units, unit) in KERNELS.items():
    path = os.path.join(src, "%sncu_%s.json" % (prefix, what))
    if not os.path.exists(path):
        continue
    d = json.load(open(path))
    name, m = next(iter(d.items()))
    res[key] = {
        "ncu_kernel": name,
        "issue_active_pct": num(m["issue_active_pct"]), "ipc": num(m["ipc"]),
        "fma_pipe_pct": num(m["fma_pipe_pct"]), "alu_pipe_pct": num(m["alu_pipe_pct"]),
        "xu_pipe_pct": num(m["xu_pipe_pct"]), "lsu_pipe_pct": num(m["lsu_pipe_pct"]),
        "registers": num(m["registers"]),
        "warp_instructions_per_unit": round(num(m["warp_instructions"]) / units, 1), "unit": unit,
        "smem_wavefronts_per_unit": round(num(m["smem_wavefronts"]) / units, 1),
        "smem_bank_conflicts_per_unit": round(num(m["smem_bank_conflicts"]) / units, 2),
        "dram_pct_of_nominal_peak": num(m["dram_pct_of_nominal"]),
        "source": "profiles/%s%s_ncu_full.txt (ncu --set full --clock-control none)" % (os.environ.get("PROFILE_PREFIX", "r2_"), key),
    }
json.dump(res, open(out, "w"), indent=1)
print("wrote", out, "with", len(res), "kernels")
