#!/usr/bin/env python3
"""Per-kernel histogram of the SASS opcodes that show how the sm_100a kernels move data and compute
(cuobjdump -sass of the built library): TMA tensor / bulk copies, mbarrier ops, FP64 vector and tensor pipe.
usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dctz_b200", "libdctz_gpu.so")
WANT = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "DMNMX", "FFMA", "FADD", "FMUL", "FMNMX", "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS",
        "RED", "SHFL", "VOTE", "REDUX", "BAR", "WARPSYNC", "F2I", "I2F", "LOP3", "IMNMX", "VIMNMX", "STL", "LDL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    print(f"# {os.path.relpath(SO, ROOT)}: architectures {arch}")
    print("# columns: kernel, total instructions, then the counts of the opcodes listed (prefix match, predicates ignored)")
    kernels = collections.OrderedDict()
    cur = None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void dctz::", "").replace("dctz::", "")
        cols = " ".join(f"{w}={c[w]}" for w in WANT if c[w])
        print(f"{short:48s} n={c['_total']:6d}  {cols}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# whole library: " + " ".join(f"{w}={tot[w]}" for w in WANT if tot[w]))
    print(f"# tcgen05 (UTC*MMA / LDTM / STTM): {sum(1 for ln in txt.splitlines() if re.search(r'UTC[A-Z]*MMA|LDTM|STTM', ln))} "
          "-- none: tcgen05 has no FP64 kind and TF32 fails the 1e-5 coefficient tolerance; the FP64 tensor route that exists (DMMA) is the "
          "k_dct64_dmma comparison kernel")


if __name__ == "__main__":
    main()
