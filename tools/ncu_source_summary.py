#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by source line:
instructions executed, stall samples and the no-instruction share.  Usage: ncu_source_summary.py src_page.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    cur_file, header = None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, samples, no_inst, text
    per_file = defaultdict(lambda: [0, 0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            i_inst = header.index("Instructions Executed")
            i_samp = header.index("# Samples")
            i_noi = header.index("stall_no_inst")
            continue
        if header is None or len(r) < len(header):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        def num(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        inst, samp, noi = num(r[i_inst]), num(r[i_samp]), num(r[i_noi])
        k = (cur_file, line)
        agg[k][0] += inst
        agg[k][1] += samp
        agg[k][2] += noi
        agg[k][3] = r[1].strip()[:110]
        per_file[cur_file][0] += inst
        per_file[cur_file][1] += samp
        per_file[cur_file][2] += noi
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total inst {tot_i}  samples {tot_s}")
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:24s} inst {100*v[0]/tot_i:5.1f}%  samples {100*v[1]/tot_s:5.1f}%  no_inst {100*v[2]/max(v[1],1):5.1f}% of its samples")
    print("--- top lines by instructions executed")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*v[0]/tot_i:5.2f}% i {100*v[1]/tot_s:5.2f}% s  {f}:{l}  {v[3]}")
    print("--- top lines by stall samples")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*v[1]/tot_s:5.2f}% s (no_inst {100*v[2]/max(v[1],1):4.0f}%) {100*v[0]/tot_i:5.2f}% i  {f}:{l}  {v[3]}")


if __name__ == "__main__":
    main()
