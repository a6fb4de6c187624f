#!/usr/bin/env python
"""Summarise an ncu report (raw page CSV) into the handful of counters DESIGN.md argues from.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "launch__shared_mem_per_block_dynamic",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = (r[i], units[i])
            if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    d.setdefault("stalls_per_issue", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 3)
        res.append(d)
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
