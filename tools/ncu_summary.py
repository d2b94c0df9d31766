#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.  usage: tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print("kernel:", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print("  %-85s %s %s" % (k, d[k], u[h.index(k)]))
        st = sorted(((float(v.replace(",", "")), k) for k, v in d.items() if "average_warps_issue_stalled" in k and "not_issued" not in k and v), reverse=True)
        print("  stalls per issue:", ", ".join("%s=%.2f" % (k.split("stalled_")[1].split("_per_")[0], v) for v, k in st[:8]))
        pipes = sorted(((float(v.replace(",", "")), k) for k, v in d.items() if k.startswith("sm__inst_executed_pipe_") and k.endswith(".sum") and v), reverse=True)
        print("  inst by pipe:", ", ".join("%s=%.3g" % (k[len("sm__inst_executed_pipe_"):-4], v) for v, k in pipes[:8]))


if __name__ == "__main__":
    main()
