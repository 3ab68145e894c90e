#!/usr/bin/env python
"""One block per kernel launch of an ncu report (`ncu -i X.ncu-rep --page raw --csv`): duration, DRAM bytes and achieved
DRAM GB/s, issue-slot / pipe utilisation, occupancy.  For reports that hold many different kernels (tools/aux_kernels.py).
Usage: ncu_multi_summary.py X.ncu-rep ["header line"] > profiles/ncu_rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio")
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def num(v, u):
    try:
        return float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)
    except ValueError:
        return None


def main():
    rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(header)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"--- {name}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
        t = num(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]) if "gpu__time_duration.sum" in col else None
        rd = num(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) if "dram__bytes_read.sum" in col else None
        wr = num(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) if "dram__bytes_write.sum" in col else None
        if t and rd is not None and wr is not None:
            print(f"    duration {t * 1e3:.3f} ms, dram read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = {(rd + wr) / t / 1e9:.0f} GB/s")
        for h in KEEP:
            if h in col and h not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
                print(f"    {h} [{units[col[h]]}] = {r[col[h]]}")
        for h in hdr:
            if "issue_stalled" in h and "per_issue_active" in h:
                v = num(r[col[h]], "")
                if v and v >= 0.5:
                    print(f"    {h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {v:.2f}")


if __name__ == "__main__":
    main()
