#!/usr/bin/env python
"""Executed-instruction mix by SASS opcode of one kernel, from an `ncu --page source --csv --print-source cuda,sass`
dump (the SASS rows carry the execution count of every instruction).

  python tools/ncu_opcode_mix.py src.csv [n_warps n_evals]
"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)


def main():
    src = sys.argv[1]
    n_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0
    n_evals = float(sys.argv[3]) if len(sys.argv) > 3 else 21.0
    mix, thr = collections.Counter(), collections.Counter()
    seen = set()
    for r in csv.reader(open(src)):
        if len(r) < 9 or r[0] != "" or not r[2].startswith("0x"):
            continue
        if r[2] in seen:          # an instruction is listed once per source line it is attributed to
            continue
        seen.add(r[2])
        toks = r[3].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        try:
            mix[op] += int(r[7]); thr[op] += int(r[8])
        except ValueError:
            pass
    tot = float(sum(mix.values()))
    k = n_warps * n_evals
    groups = {"fp32 arithmetic": ("FFMA", "FMUL", "FADD", "FFMA2", "FMNMX", "FSEL", "FSETP", "FCHK", "MUFU", "FSET"),
              "integer / address": ("IMAD", "IADD3", "IADD", "LEA", "LOP3", "SHF", "ISETP", "VIADD", "IABS", "PRMT", "SEL", "I2F", "F2I", "I2FP", "F2FP", "POPC", "FLO", "BREV", "IMNMX", "VIMNMX", "SGXT", "LOP", "PLOP3", "P2R", "R2P", "ULOP3", "UIADD3", "UMOV", "USHF", "ULEA", "UIMAD", "UISETP", "R2UR", "S2R", "S2UR", "CS2R"),
              "moves": ("MOV", "HFMA2", "IMAD.MOV"),
              "shared / local / global memory": ("LDS", "STS", "LDL", "STL", "LDG", "STG", "LD", "ST", "LDC", "ULDC", "LDSM", "ATOMS", "ATOMG", "RED", "MEMBAR", "CCTL"),
              "shuffles / votes": ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"),
              "control": ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "NOP", "BMOV", "BREAK", "YIELD", "NANOSLEEP", "DEPBAR", "ERRBAR", "JMP", "BRX")}
    gsum = collections.Counter()
    for op, n in mix.items():
        for g, ops in groups.items():
            if op in ops:
                gsum[g] += n
                break
        else:
            gsum["other"] += n
    print("total %.0f warp-instructions (%.0f per evaluation and warp)" % (tot, tot / k))
    for g, n in gsum.most_common():
        print("  %-32s %5.1f %%  %7.1f per evaluation and warp" % (g, 100 * n / tot, n / k))
    print()
    for op, n in mix.most_common(28):
        print("  %-10s %5.1f %%  %7.1f per eval/warp   threads/inst %.1f" % (op, 100 * n / tot, n / k, thr[op] / max(n, 1)))


if __name__ == "__main__":
    main()
