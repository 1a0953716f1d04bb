#!/usr/bin/env python
"""SASS instructions (with their execution counts) attributed to given source lines of one file, from an
`ncu --page source --csv --print-source cuda,sass` dump.

  python tools/ncu_sass_of_line.py src.csv bio_coop_spatial.cuh 419 420 [n_warps]
"""
import csv
import os
import sys

csv.field_size_limit(10 ** 9)


def main():
    src, fname = sys.argv[1], sys.argv[2]
    lines = set(int(a) for a in sys.argv[3:] if a.isdigit() and int(a) < 100000)
    n_warps = 8192.0
    cur, want = None, False
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if r[0] in ("Line No", "Function Name"):
            continue
        if r[0].strip().isdigit():
            want = cur == fname and int(r[0]) in lines
            if want:
                print("---- %s:%s  %s" % (cur, r[0], r[1][:110]))
            continue
        if want and r[0] == "" and len(r) > 7:
            try:
                print("    %8.2f  %s" % (int(r[7]) / n_warps / 21.0, r[3].strip()))
            except ValueError:
                pass


if __name__ == "__main__":
    main()
