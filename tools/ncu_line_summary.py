#!/usr/bin/env python3
"""Per CUDA source line: stall samples, executed warp instructions, shared-memory wavefronts of one kernel of an ncu report,
by joining `ncu --page source --csv` (SASS rows with addresses) with the line table of the SAME build (nvdisasm -g).
usage: ncu_line_summary.py report.ncu-rep libdctz_gpu.so <kernel regex> <mangled-name substring> [launch-skip] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def line_table(so, mangled):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(txt) if ".section\t.text." in l and mangled in l][0]
    cur, tab = None, {}
    for ln in txt[start + 1:]:
        if ".section" in ln:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            tab[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return tab


def main():
    rep, so, kre, mangled = sys.argv[1:5]
    skip = sys.argv[5] if len(sys.argv) > 5 else "0"
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
    tab = line_table(so, mangled)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    iA, iS, iN, iX = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    iW = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "")
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0.0, 0])
    for r in rows:
        if len(r) <= iX or not r[iN].isdigit():
            continue
        a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
        if base is None:
            base = a
        key = tab.get(a - base, (None, ""))[0]
        g = agg[key]
        g[0] += int(r[iN]); g[1] += int(r[iX] or 0); g[3] += 1
        if iW is not None and r[iW]:
            g[2] += float(r[iW])
    tot = sum(g[0] for g in agg.values()); totx = sum(g[1] for g in agg.values())
    print(f"samples {tot}, executed warp instructions {totx}")
    src_cache = {}
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            for root in ("dctz_b200/csrc",):
                p = os.path.join(root, key[0])
                if os.path.exists(p):
                    src_cache.setdefault(p, open(p).read().split("\n"))
                    text = src_cache[p][key[1] - 1].strip()[:110]
        print(f"{100 * g[0] / max(tot, 1):5.1f}% samples  {100 * g[1] / max(totx, 1):5.1f}% instr ({g[3]:4d} static) smem wavefronts {g[2]:.3g}  {key}: {text}")


if __name__ == "__main__":
    main()
