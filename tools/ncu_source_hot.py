#!/usr/bin/env python3
"""Hot source lines of one kernel from a full ncu report: samples and instructions per CUDA source line.
usage: ncu_source_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections

def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur = None; hdr = None
    rows = []
    for r in csv.reader(out.splitlines()):
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or r[0] == "": continue
        d = dict(zip(hdr, r))
        # duplicate column name "Source": take by position
        rows.append((cur, int(r[0]), r[1].strip(), d))
    tot_s = sum(num(d["# Samples"]) for _, _, _, d in rows) or 1
    tot_i = sum(num(d["Instructions Executed"]) for _, _, _, d in rows) or 1
    per_file = collections.Counter()
    for f, _, _, d in rows: per_file[f] += num(d["# Samples"])
    print("samples per file:", {k: "%.1f%%" % (100.0 * v / tot_s) for k, v in per_file.most_common()})
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st_tot = collections.Counter()
    for _, _, _, d in rows:
        for s in stalls: st_tot[s] += num(d[s])
    print("stall mix:", {k: "%.1f%%" % (100.0 * v / tot_s) for k, v in st_tot.most_common(8)})
    rows.sort(key=lambda x: -num(x[3]["# Samples"]))
    print("%-14s %5s %7s %7s %5s  %-28s %s" % ("file", "line", "smp%", "inst%", "thr", "top stalls", "source"))
    for f, ln, src, d in rows[:top]:
        s = num(d["# Samples"])
        ts = sorted(((num(d[k]), k[6:]) for k in stalls), reverse=True)[:2]
        print("%-14s %5d %6.2f%% %6.2f%% %5s  %-28s %s" % (f, ln, 100.0 * s / tot_s, 100.0 * num(d["Instructions Executed"]) / tot_i,
              d["Avg. Threads Executed"], ",".join("%s:%d" % (k, v) for v, k in ts if v), src[:90]))

main()
