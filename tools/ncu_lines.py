#!/usr/bin/env python
"""Warp-instructions per evaluation and warp of every source line of one file, from an
`ncu --page source --csv --print-source cuda,sass` dump (see tools/ncu_phase_report.py).

  python tools/ncu_lines.py src.csv bio_coop_spatial.cuh 190 300 [n_warps n_evals]
"""
import collections
import csv
import os
import sys

csv.field_size_limit(10 ** 9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    n_warps = float(sys.argv[5]) if len(sys.argv) > 5 else 8192.0
    n_evals = float(sys.argv[6]) if len(sys.argv) > 6 else 21.0
    cur, hdr = None, None
    ins, smp, wav, exc = (collections.Counter() for _ in range(4))
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1]); hdr = None
            continue
        if r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            iW, iE = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
            continue
        if cur != fname or hdr is None or not r[0].strip().isdigit():
            continue
        try:
            ln = int(r[0]); ins[ln] += int(r[iI]); smp[ln] += int(r[iS])
            wav[ln] += int(r[iW] or 0); exc[ln] += int(r[iE] or 0)
        except (ValueError, IndexError):
            continue
    text = open(os.path.join(ROOT, "bioimitation_gym_b200", "csrc", fname)).read().split("\n")
    k = n_warps * n_evals
    tot = 0.0
    for ln in range(lo, hi + 1):
        if ins[ln]:
            tot += ins[ln] / k
            print("%4d %7.1f inst %6d smp  wavefronts %6.1f (excess %5.1f)  %s" % (ln, ins[ln] / k, smp[ln], wav[ln] / k, exc[ln] / k, text[ln - 1][:100]))
    print("lines %d..%d: %.1f warp-instructions per evaluation and warp" % (lo, hi, tot))


if __name__ == "__main__":
    main()
