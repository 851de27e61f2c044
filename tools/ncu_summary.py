#!/usr/bin/env python3
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel, and optionally the
key metrics of a full .ncu-rep.  usage: ncu_summary.py launches.csv [report.ncu-rep]"""
import collections, csv, subprocess, sys

def launch_summary(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt, mx = collections.Counter(), collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        tot[name] += v; cnt[name] += 1; mx[name] = max(mx[name], v)
    T = sum(tot.values())
    print("launches %d, kernel time %.3f ms (cold-cache, serialised: compare SHARES)" % (sum(cnt.values()), T / 1e6))
    for k, v in tot.most_common():
        print("%-20s n=%5d total=%10.1f us share=%5.1f%% avg=%8.2f us max=%9.1f us" % (k, cnt[k], v / 1e3, 100 * v / T, v / 1e3 / cnt[k], mx[k] / 1e3))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

def report_summary(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernels:", [r[idx["Kernel Name"]].split("(")[0] for r in rows[2:]])
    for w in WANT:
        if w in idx:
            print("%-82s %-16s %s" % (w, units[idx[w]], [r[idx[w]] for r in rows[2:]]))

if __name__ == "__main__":
    launch_summary(sys.argv[1])
    if len(sys.argv) > 2:
        print()
        report_summary(sys.argv[2])
