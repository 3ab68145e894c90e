#!/usr/bin/env python
"""Static code size of one kernel by source line, from `nvdisasm --print-line-info X.cubin`.
Usage: sass_static_summary.py dis.txt <kernel-substring> [top_n]"""
import re
import sys
from collections import Counter


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cnt, inside, cur = Counter(), False, None
    for line in open(path):
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            cnt[cur] += 1
    total = sum(cnt.values())
    print("total instructions", total, "=", total * 16 // 1024, "KB")
    files = Counter()
    for (f, l), c in cnt.items():
        files[f] += c
    print(files.most_common())
    src = {}
    for (f, l), c in cnt.most_common(top):
        if f not in src:
            try:
                src[f] = open("/root/repo/era_zk_evm_b200/csrc/" + f).read().split("\n")
            except OSError:
                src[f] = []
        text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
        print(f"{c:5d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
