#!/usr/bin/env python
"""Per-kernel extract of an ncu report (the table format of profiles/r2_ncu_kernels.csv: one row per metric, one column per
captured launch):

    python tools/ncu_extract.py report.ncu-rep "header comment" > profiles/<name>.csv
"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main():
    rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled_") and
                                                          h.endswith("_per_issue_active.ratio") and "not_issued" not in h)]
    out = io.StringIO()
    w = csv.writer(out)
    print(f"# ncu --set full --clock-control none --import-source on: {note}")
    print("# metric, unit, value per captured launch")
    for i in cols:
        w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    sys.stdout.write(out.getvalue())


if __name__ == "__main__":
    main()
