#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump of the
cooperative step kernel into instruction / stall-sample shares per phase
(phases are the `// ---- phase` markers of the kernel sources) and per helper.

  ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_phase_report.py src.csv [n_warps n_evals] [--lines K]

The CSV holds one section per source file ("File Path" header row); lines of
bio_coop.cuh / bio_coop_planar.cuh are binned by the nearest phase marker above
them, lines of the other files by file (helpers).  --lines K also prints the K
hottest source lines.  The copy of the sources the profile was taken from must be
the one in the tree (line numbers).
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bioimitation_gym_b200", "csrc")
MARK = re.compile(r"// ---- (phase \w+|full evaluation|action pre|integrate one|reward|termination|write back|solve \w+)")


def phase_marks(path):
    marks = []
    if not os.path.exists(path):
        return marks
    for i, ln in enumerate(open(path), 1):
        m = MARK.search(ln)
        if m:
            marks.append((i, m.group(1)))
        m = re.match(r"^(?:__device__|__global__|BIO_DEV)[^(]*?(\w+)\(", ln)
        if m:
            marks.append((i, "fn:" + m.group(1)))
    return marks


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_lines = 0
    if "--lines" in sys.argv:
        n_lines = int(sys.argv[sys.argv.index("--lines") + 1])
        args = [a for a in args if a != str(n_lines)] if str(n_lines) in args[1:] else args
    src = args[0]
    n_warps = float(args[1]) if len(args) > 1 else 2048.0
    n_evals = float(args[2]) if len(args) > 2 else 21.0
    agg, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    lines = collections.Counter()
    line_smp = collections.Counter()
    line_text = {}
    cur_file, marks, hdr = None, [], None
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            marks = phase_marks(os.path.join(CSRC, cur_file))
            continue
        if r[0] == "Line No":
            hdr = r
            iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or not r[0].strip().isdigit() or len(r) <= iT:
            continue
        try:
            ins, t, s = int(r[iI]), int(r[iT]), int(r[iS])
        except ValueError:
            continue
        ln = int(r[0])
        key = None
        for a, name in marks:
            if a <= ln:
                key = name
        if cur_file in ("bio_coop.cuh", "bio_coop_planar.cuh"):
            key = "%s" % (key or "top")
        else:
            key = "%s:%s" % (cur_file, key or "-")
        agg[key] += ins
        thr[key] += t
        smp[key] += s
        lines[(cur_file, ln)] += ins
        line_smp[(cur_file, ln)] += s
        line_text[(cur_file, ln)] = r[1].strip()
    tot, tots = sum(agg.values()), sum(smp.values())
    print("total warp-instructions %d (%.0f per eval per warp), samples %d" % (tot, tot / n_warps / n_evals, tots))
    for k, v in agg.most_common(45):
        print("%-40s %5.1f%% inst %5.1f%% samples %7.0f instr/eval/warp  thr/inst %4.1f" % (
            k, 100.0 * v / tot, 100.0 * smp[k] / max(tots, 1), v / n_warps / n_evals, thr[k] / max(v, 1)))
    if n_lines:
        print("\nhottest source lines (by stall samples)")
        for (f, ln), s in line_smp.most_common(n_lines):
            print("%5.2f%% smp %5.2f%% inst  %s:%d  %s" % (100.0 * s / max(tots, 1), 100.0 * lines[(f, ln)] / tot, f, ln,
                                                       line_text[(f, ln)][:110]))


if __name__ == "__main__":
    main()
