#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump of the
cooperative step kernel into instruction / stall-sample shares per phase
(phases are the `// ---- phase` markers of bio_coop.cuh) and per helper.

  ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_phase_report.py src.csv [n_warps n_evals]
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def phase_ranges():
    path = os.path.join(ROOT, "bioimitation_gym_b200", "csrc", "bio_coop.cuh")
    marks = []
    for i, ln in enumerate(open(path), 1):
        m = re.search(r"// ---- (phase \w+|full evaluation|action pre|integrate one|reward|termination|write back)", ln)
        if m:
            marks.append((i, m.group(1)))
        if re.match(r"^(template|__global__)", ln):
            marks.append((i, "fn@%d" % i))
    marks.sort()
    return marks


def main():
    src = sys.argv[1]
    n_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
    n_evals = float(sys.argv[3]) if len(sys.argv) > 3 else 21.0
    rows = list(csv.reader(open(src)))
    for hi, r in enumerate(rows):
        if "Line No" in r and "Instructions Executed" in r:
            break
    hdr = rows[hi]
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    marks = phase_ranges()
    coop_lines = {}
    for ln in open(os.path.join(ROOT, "bioimitation_gym_b200", "csrc", "bio_coop.cuh")):
        pass
    coop_src = [l.rstrip("\n").strip() for l in open(os.path.join(ROOT, "bioimitation_gym_b200", "csrc", "bio_coop.cuh"))]
    agg, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= iI or not r[0].strip().isdigit():
            continue
        try:
            ins, t, s = int(r[iI]), int(r[iT]), int(r[iS])
        except ValueError:
            continue
        ln, text = int(r[0]), r[1].strip()
        key = None
        if ln <= len(coop_src) and coop_src[ln - 1][:60] == text[:60]:
            for a, name in marks:
                if a <= ln:
                    key = name
        if key is None:
            m = re.match(r".*?(cross3|matvec3|dot3|clampv|func_eval|curve_eval|step5|sincos|sqrt|__syncwarp|bar_warp)", text)
            key = "helper:" + (m.group(1) if m else "other@%d" % ln if ins > 3e6 else "misc")
        agg[key] += ins
        thr[key] += t
        smp[key] += s
    tot, tots = sum(agg.values()), sum(smp.values())
    print("total warp-instructions %d (%.0f per eval per warp), samples %d" % (tot, tot / n_warps / n_evals, tots))
    for k, v in agg.most_common(40):
        print("%-28s %5.1f%% inst %5.1f%% samples %7.0f instr/eval/warp  thr/inst %4.1f" % (
            k, 100.0 * v / tot, 100.0 * smp[k] / max(tots, 1), v / n_warps / n_evals, thr[k] / max(v, 1)))


if __name__ == "__main__":
    main()
