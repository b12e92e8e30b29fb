#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt [kernel-substring]   (also writes profiles/r1_xxx.json: the headline metrics)"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active_mem_shared.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
JSON_KEYS = {"gpu__time_duration.sum": "ncu_time_us", "smsp__inst_executed.sum": "warp_instructions",
             "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__inst_executed.avg.per_cycle_active": "ipc",
             "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
             "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_nominal", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
             "launch__registers_per_thread": "registers"}

def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    js = {}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % rep)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if sub and sub not in d.get("Kernel Name", ""):
                continue
            f.write("\n## launch %s: %s  grid %s block %s\n" % (d.get("ID"), d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
            for k in KEYS:
                if k in d:
                    f.write("%-70s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            js.setdefault(d.get("Kernel Name", "?"), {k2: (d[k1] + " " + units[hdr.index(k1)]).strip() for k1, k2 in JSON_KEYS.items() if k1 in d})
            stalls = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")]
            for v, k in sorted(stalls, reverse=True)[:8]:
                f.write("%-70s %.3f\n" % (k.replace("smsp__average_warps_issue_stalled_", "stall/issue: ").replace("_per_issue_active.ratio", ""), v))
    print(open(out).read())
    if out.endswith(".txt"):
        import json
        with open(out[:-4] + ".json", "w") as f:
            json.dump(js, f, indent=1)

if __name__ == "__main__":
    main()
