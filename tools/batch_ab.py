import os, sys, time, json, re
sys.path.insert(0, "/root/repo")
import ries_b200, bench
g = json.load(open("tests/golden/batch_l4.json"))
flags = g["flags"] + ["--no-slow-messages"]
ts = bench.targets_for(256)
ries_b200.batch(ts[:16], flags, workers=8)
rates = []
for rep in range(int(os.environ.get("REPS", "4"))):
    t0 = time.perf_counter(); got = ries_b200.batch(ts, flags, workers=int(os.environ.get("W", "8"))); dt = time.perf_counter() - t0
    rates.append(round(len(ts) / dt, 1))
print("rates", rates)
want = dict(zip(g["targets"], g["reports"]))
bad = [t for t, (c, x) in zip(ts, got) if t in want and bench.normalise(x) + "exit=%d\n" % bench.shown_status(c) != want[t]]
print(os.environ.get("RIES_BATCH_NO_SHARE", "share"), "W", os.environ.get("W", "8"), "targets/s %.1f" % (len(ts) / dt), "mismatch", len(bad), bad[:3], ries_b200.batch_shared_stats())
if bad:
    t = bad[0]; c, x = got[ts.index(t)]
    a = (bench.normalise(x) + "exit=%d\n" % bench.shown_status(c)).splitlines(); b = want[t].splitlines()
    for u, v in zip(a, b):
        if u != v: print("GOT ", u); print("WANT", v)
