#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` per enclosing device function of the source files:
instructions executed and stall samples per function.  Usage: ncu_region_summary.py src_page.csv <git-rev-of-sources|-> """
import csv, re, subprocess, sys, os
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "era_zk_evm_b200", "csrc")

def load(fname, rev):
    if rev and rev != "-":
        return subprocess.check_output(["git", "-C", ROOT, "show", f"{rev}:era_zk_evm_b200/csrc/{fname}"], text=True).split("\n")
    return open(os.path.join(CSRC, fname)).read().split("\n")

def func_map(lines):
    cur, out = "<top>", []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b([A-Za-z_][A-Za-z0-9_:]*)\s*\(")
    for ln in lines:
        m = pat.match(ln)
        if m and not ln.rstrip().endswith(";"):
            cur = m.group(1)
        out.append(cur)
    return out

def main():
    path, rev = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "-")
    rows = list(csv.reader(open(path, newline="")))
    maps = {}
    agg = defaultdict(lambda: [0, 0])
    cur_file = header = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            header = r; i_inst = header.index("Instructions Executed"); i_samp = header.index("# Samples"); continue
        if header is None or len(r) < len(header): continue
        try: line = int(r[0])
        except ValueError: continue
        def num(x):
            try: return int(float(x))
            except ValueError: return 0
        key = cur_file
        if cur_file and cur_file.endswith((".cuh", ".cu")):
            if cur_file not in maps:
                try: maps[cur_file] = func_map(load(cur_file, rev))
                except Exception: maps[cur_file] = []
            fm = maps[cur_file]
            key = f"{cur_file}:{fm[line - 1] if 0 < line <= len(fm) else '?'}"
        agg[key][0] += num(r[i_inst]); agg[key][1] += num(r[i_samp])
    ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
    print(f"total inst {ti} samples {ts}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{100*v[0]/ti:6.2f}% inst {100*v[1]/ts:6.2f}% samples  {k}")
main()
