#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: per-opcode static count / executed / stall samples,
and the top stalled instructions.  usage: ncu_src_summary.py file.csv [top_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    ops, samp, execd = collections.Counter(), collections.Counter(), collections.Counter()
    insts = []
    for r in rows:
        if len(r) <= iEx or not r[iSamp].isdigit():
            continue
        parts = r[iS].strip().split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.split(".")[0]
        ops[op] += 1
        samp[op] += int(r[iSamp])
        execd[op] += int(r[iEx] or 0)
        insts.append((int(r[iSamp]), len(insts), r[iS].strip(), int(r[iEx] or 0)))
    tot, totex = sum(samp.values()), sum(execd.values())
    print("static instrs", len(insts), "samples", tot, "executed warp-instr", totex)
    for op, c in samp.most_common(top_n):
        print(f"{op:10s} static {ops[op]:6d} samples {c:7d} {100 * c / max(tot,1):5.1f}%  exec {execd[op]:10d} {100 * execd[op] / max(totex,1):5.1f}%")
    print("--- top instructions by samples")
    for s, i, src, ex in sorted(insts, reverse=True)[:top_n]:
        print(f"{s:6d} @{i:6d} exec {ex:8d}  {src}")


if __name__ == "__main__":
    main()
