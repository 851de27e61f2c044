#!/usr/bin/env python3
"""A/B of kernel-tuning variants on the GPU box: for every _ab/*.so, a fresh process runs the full -l6 search
(3 warm-ups + 5 timed), checks the report against the golden, and prints wall / per-stage device times.
usage: tools/ab_l6.py [--env NAME:K=V,K=V ...]   (driver; every --env is one more run of the shipped library with
those environment variables)      tools/ab_l6.py --one   (worker)"""
import glob, os, re, shutil, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

def one():
    import ries_b200
    gold = open(os.path.join(ROOT, "tests", "golden", "l6full.txt")).read()
    args = ["2.5063141592653589", "-l6", "-F3", "--no-slow-messages"]
    walls, stages = [], []
    ok = True
    for rep in range(8):
        t0 = time.perf_counter(); code, text = ries_b200.call_main(args); dt = time.perf_counter() - t0
        st = ries_b200.last_stats()
        if rep >= 3:
            walls.append(dt * 1e3); stages.append(st["stage_ms"])
        got = re.sub(r"Time: [0-9.]+", "Time: X", text) + "exit=%d\n" % (code % 256)
        ok = ok and (got == gold)
    med = sorted(walls)[len(walls) // 2]
    avg = [sum(s[i] for s in stages) / len(stages) for i in range(4)]
    print("%-28s parity=%s wall med %.1f min %.1f ms | K1 %.2f sort %.2f merge %.2f refine %.2f" % (
        os.environ.get("AB_NAME", "?"), "OK" if ok else "FAIL", med, min(walls), avg[0], avg[1], avg[2], avg[3]), flush=True)

if "--one" in sys.argv:
    one()
else:
    lib = os.path.join(ROOT, "ries_b200", "libries_b200.so")
    shutil.copy(lib, lib + ".keep")
    envs = []
    for i, a in enumerate(sys.argv):
        if a == "--env" and i + 1 < len(sys.argv):
            name, _, kv = sys.argv[i + 1].partition(":")
            envs.append((name, dict(x.split("=", 1) for x in kv.split(",") if x)))
    try:
        for v in [lib + ".keep"] + sorted(glob.glob(os.path.join(ROOT, "_ab", "*.so"))):
            shutil.copy(v, lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, AB_NAME=os.path.basename(v)), timeout=300)
        shutil.copy(lib + ".keep", lib)
        for name, extra in envs:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, AB_NAME=name, **extra), timeout=300)
    finally:
        shutil.copy(lib + ".keep", lib); os.remove(lib + ".keep")
