#!/usr/bin/env python
"""Writes the text summary of one kernel from an ncu report: `ncu -i X.ncu-rep --page raw --csv` filtered to the metrics
the roofline discussion uses (duration, DRAM traffic, issue slots, pipe utilisation, stall reasons, occupancy limits).
Usage: ncu_summarize.py X.ncu-rep "header line" > profiles/ncu_rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__icc_request_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum")


def main():
    rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(header)
    for h, u, v in sorted(zip(hdr, units, vals)):
        if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h):
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
