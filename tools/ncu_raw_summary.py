#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of counters the roofline discussion
uses (one block per kernel launch).  usage: ncu_raw_summary.py raw.csv > summary.txt"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("==", name.split("(")[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:70s} {r[i]:>18s} {units[i]}")
        rd, wr, t = (float(r[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
        wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
        t *= {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
        print(f"   {'dram traffic (read+write) / duration':70s} {(rd + wr) / t / 1e9:18.1f} GB/s   traffic {(rd + wr) / 1e9:.4f} GB")
        idx = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        st = sorted(((float(r[i]), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i in idx if r[i]), reverse=True)[:7]
        print("   top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st))


if __name__ == "__main__":
    main()
