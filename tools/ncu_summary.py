#!/usr/bin/env python
"""Summarise an .ncu-rep (read on a box without a GPU): per-launch key metrics + instruction mix + top stall lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of %s (captured with --set full --clock-control none)" % rep)
    for r in rows[2:]:
        print("\n## launch: %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("%-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 3:
        h = src[1]
        ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        sec = []
        for r in src[2:]:
            if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
                break
            sec.append(r)
        tot = sum(int(r[ie]) for r in sec)
        ts = max(sum(int(r[iss]) for r in sec), 1)
        print("\n## first launch: %d SASS instructions, %d warp-level instructions executed" % (len(sec), tot))
        ops = collections.Counter()
        for r in sec:
            t = r[ia].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += int(r[ie])
        print("instruction mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in ops.most_common(12)))
        print("top stall samples (share, executed, SASS):")
        for r in sorted(sec, key=lambda r: -int(r[iss]))[:8]:
            print("  %5.1f%%  %9s  %s" % (100.0 * int(r[iss]) / ts, r[ie], r[ia][:90]))


if __name__ == "__main__":
    main()
