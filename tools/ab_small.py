#!/usr/bin/env python3
"""A/B on the GPU box of the latency-bound configurations (default level, -l4, -l5 -i): for every _ab/*.so and
every --env NAME:K=V,... a fresh process runs each search 3 + 9 times and prints the median wall time and a
hash of the report (equal hashes = same bytes).   usage: tools/ab_small.py [--env NAME:K=V,K=V ...]"""
import glob, hashlib, os, re, shutil, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
CASES = [("default", ["3.14159265358979", "-F3"]), ("l4", ["2.5063", "-l4", "-F3"]), ("l5i", ["143", "-i", "-l5", "-F3"])]

def one():
    import ries_b200
    out = []
    for name, a in CASES:
        walls = []; h = ""
        for rep in range(12):
            t0 = time.perf_counter(); code, text = ries_b200.call_main(a + ["--no-slow-messages"]); dt = time.perf_counter() - t0
            if rep >= 3: walls.append(dt * 1e3)
            h = hashlib.md5(re.sub(r"Time: [0-9.]+", "Time: X", text).encode()).hexdigest()[:8]
        out.append("%s %.2f ms (min %.2f) %s" % (name, sorted(walls)[len(walls) // 2], min(walls), h))
    print("%-24s %s" % (os.environ.get("AB_NAME", "?"), " | ".join(out)), flush=True)

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
