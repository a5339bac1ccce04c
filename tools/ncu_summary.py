#!/usr/bin/env python
"""Text summary of an .ncu-rep (read here, on the CPU box, with `ncu -i`): per kernel the duration,
instruction count, lane efficiency, issue utilisation, occupancy, L1/L2 hit rates, DRAM bytes and the
top warp-stall reasons; optionally the hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--lines N] > profiles/x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (elapsed)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "blocks / SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "blocks / SM (shared-memory limit)"),
    ("smsp__cycles_active.avg", "SM sub-partition cycles active"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def lines(rep, kernel_id, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--launch-skip", str(kernel_id), "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif r and r[0] and r[0].isdigit() and hdr:
            def num(v):
                try:
                    return float(v)
                except ValueError:
                    return 0.0
            i, s = num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])
            a = agg.get((cur, int(r[0])), (0.0, 0.0, ""))
            agg[(cur, int(r[0]))] = (a[0] + i, a[1] + s, r[1].strip()[:100])
    ti = sum(v[0] for v in agg.values()) or 1.0
    ts = sum(v[1] for v in agg.values()) or 1.0
    print("  hottest source lines (share of executed warp instructions / of stall samples):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("    %-18s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / ti, 100 * v[1] / ts, v[2]))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
    hdr, units, rows = raw(rep)
    print("ncu summary of %s  (ncu --set full --clock-control none; cold-cache, serialised replays)" % rep.split("/")[-1])
    for n, r in enumerate(rows):
        d = dict(zip(hdr, r))
        print("\n== %s   grid %s block %s" % (d["Kernel Name"][:90], d.get("Grid Size"), d.get("Block Size")))
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("  %-40s %14s %s" % (label, d[k], units[hdr.index(k)]))
        st = {}
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v:
                st[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
        print("  warp stalls per issued instruction: " + ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
        if top:
            lines(rep, n, top)


if __name__ == "__main__":
    main()
