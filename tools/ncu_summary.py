#!/usr/bin/env python
"""Turn ncu output into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/rNN_x_launches.txt
  python tools/ncu_summary.py full     gpurun_out/x.ncu-rep       > profiles/rNN_x_full.txt
  python tools/ncu_summary.py stalls   gpurun_out/x.ncu-rep [N]   > profiles/rNN_x_stalls.txt   (top-N source lines)
"""
import collections
import csv
import io
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel totals of gpu__time_duration.sum (ncu, cold-cache, serialised: compare SHARES, not absolutes)")
    print("# source: %s ; %d launches, %.3f ms in total" % (path, sum(a[0] for a in agg.values()), tot / 1e6))
    print("%-44s %7s %14s %12s %8s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s %7d %14.1f %12.2f %8.4f" % (k[:44], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[1] / tot))


def full(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[hdr.index("Kernel Name")])
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-84s %16s %s" % (k, r[i], units[i]))


def stalls(rep, top=25):
    rows = ncu_csv(rep, "source")
    # find the header row of the source table
    for hi, r in enumerate(rows):
        if "Source" in r and any("Sampling" in c for c in r):
            break
    else:
        print("no source page")
        return
    hdr = rows[hi]
    si = hdr.index("Source")
    samp = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
    ci = samp[0] if samp else None
    body = []
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or ci is None:
            continue
        try:
            body.append((float(r[ci] or 0), r[si].strip()))
        except ValueError:
            pass
    tot = sum(b[0] for b in body) or 1.0
    print("# top source lines by warp-stall samples (%s), total %d" % (hdr[ci], tot))
    for v, s in sorted(body, key=lambda x: -x[0])[:top]:
        print("%8d %6.2f%%  %s" % (v, 100 * v / tot, s[:150]))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2])
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
