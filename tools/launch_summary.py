#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and
share of device time per kernel (cold-cache, serialised times: compare shares)."""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")[:70]
        tot[name] += float(r[-1])
        cnt[name] += 1
    s = sum(tot.values())
    print("%d launches, %.1f us total" % (len(rows), s / 1e3))
    for k, v in tot.most_common():
        print("%6.2f%%  %9.1f us  %4d x  %s" % (100 * v / s, v / 1e3, cnt[k], k))
    # the step loop of bench.py launches the step kernel (timed) and, between steps and untimed, the L2 flush
    # (FillFunctor<unsigned char>); everything else is set-up (reset, equilibrium table, action pool) or the
    # FP32-peak microbenchmark that runs after the loop
    step = sum(v for k, v in tot.items() if "step_kernel" in k)
    own = sum(v for k, v in tot.items() if k.startswith("bio::"))
    print("share of the step kernel in the launches of this library (bio::*): %.1f %%" % (100 * step / max(own, 1e-9)))
    print("timed region of bench.py = the step-kernel launches only (one per control step)")


if __name__ == "__main__":
    main()
