#!/usr/bin/env python
"""profiles/r02_jac_rx_profile.json from an `ncu --set full` capture of ONE jac_rx_kernel launch: DRAM bytes per Jacobian system-step and
the pipe figures bench.py quotes next to the canonical roofline fraction.  The source hash of the library is recorded; bench.py refuses
a profile taken from other sources.  usage: tools/ncu_profile_json.py report.ncu-rep steps_per_system_in_launch out.json"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nbodygradient.jl_b200"))
from nbgrad.build import source_hash  # noqa: E402


def main():
    rep, steps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    d = dict(zip(h, rows[2]))
    unit = dict(zip(h, u))

    def val(k):
        v = float(d[k].replace(",", ""))
        s = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        return v * s.get(unit[k], 1.0)

    grid = int(float(d["launch__grid_size"].replace(",", "")))
    jsteps = grid * steps
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    n = 8
    P = n * (n - 1) // 2
    out_j = {"source": "ncu --set full --clock-control none, one jac_rx_kernel launch: %d systems x %d steps" % (grid, steps),
             "source_hash": source_hash(), "kernel": d.get("Kernel Name", "")[:80],
             "registers_per_thread": int(float(d["launch__registers_per_thread"])),
             "dram_bytes_read": rd, "dram_bytes_write": wr, "jacobian_steps_in_launch": jsteps,
             "dram_bytes_per_jacobian_step": (rd + wr) / jsteps,
             "algorithmic_bytes_per_jacobian_step": (2 * P * 64 + 12 * n * n) * 8 + 2 * 2 * 48 * 56 * 8 / steps,
             "duration_ms": val("gpu__time_duration.sum") / (1e6 if unit["gpu__time_duration.sum"] in ("ns", "nsecond") else (1e3 if unit["gpu__time_duration.sum"] in ("us", "usecond") else 1.0)),
             "duration_unit_reported": unit["gpu__time_duration.sum"],
             "fp64_pipe_cycles_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
             "smem_lsu_wavefronts_pct": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
             "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
             "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"])}
    json.dump(out_j, open(out, "w"), indent=1)
    print(json.dumps(out_j))


if __name__ == "__main__":
    main()
