#!/usr/bin/env python
"""Shared-memory wavefronts along the kernel text, from an `ncu --page source --csv --print-source cuda,sass` dump:
every LDS / STS instruction once (by address), binned by code region, with the source lines it is attributed to.

  python tools/ncu_smem_map.py src.csv n_warps [n_evals] [--list]
"""
import collections
import csv
import os
import sys

csv.field_size_limit(10 ** 9)


def main():
    src = sys.argv[1]
    n_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
    n_evals = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 21.0
    rows, cur, hdr, line = [], None, None, None
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = os.path.basename(r[1]), None
            continue
        if r[0] == "Line No":
            hdr = r
            iW, iE, iI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0].strip().isdigit():
            line = (cur, int(r[0]))
            continue
        if r[0] == "" and len(r) > iE and r[2].startswith("0x"):
            try:
                rows.append((int(r[2], 16), r[3].strip(), int(r[iI] or 0), int(r[iW] or 0), int(r[iE] or 0), line))
            except ValueError:
                pass
    rows.sort()
    k = n_warps * n_evals
    base, seen = rows[0][0], set()
    reg, exc, lines = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    tot = tote = 0.0
    for a, txt, n, w, e, l in rows:
        if a in seen or w == 0:
            continue
        seen.add(a)
        tot += w / k
        tote += e / k
        b = (a - base) // 0x800
        reg[b] += w / k
        exc[b] += e / k
        lines[b]["%s:%d" % (l[0].replace("bio_", "").replace(".cuh", ""), l[1])] += w / k
        if "--list" in sys.argv:
            print("%6x %-56s n %6.2f wav %6.2f exc %6.2f %s:%d" % (a - base, txt[:56], n / k, w / k, e / k, l[0], l[1]))
    print("shared-memory wavefronts per evaluation and warp: %.1f (excess %.1f)" % (tot, tote))
    for b in sorted(reg):
        if reg[b] >= 1.0:
            print("%6x..  %6.1f (exc %5.1f)  %s" % (b * 0x800, reg[b], exc[b], ", ".join("%s %.0f" % x for x in lines[b].most_common(5))))


if __name__ == "__main__":
    main()
