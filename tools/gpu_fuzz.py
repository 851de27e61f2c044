#!/usr/bin/env python3
"""Differential fuzz on a GPU box: random command lines (tools/fuzz_cli.py) through the product binary and
through the prebuilt unmodified reference oracle/_ref/ries.  Cases the reference needs more than 20 s for are
skipped; a product run of more than 120 s is a failure.  usage: tools/gpu_fuzz.py [n] [seed]
(FUZZ_LEVELS="2 3" selects the -l values)"""
import os, re, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_cli
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ref = os.path.join(ROOT, "oracle", "_ref", "ries")
exe = os.path.join(ROOT, "ries_b200", "bin", "ries")
env = dict(os.environ, HOME="/nonexistent")

def run(binary, args, limit):
    try:
        r = subprocess.run([binary] + args + ["--no-slow-messages"], capture_output=True, text=True, env=env, timeout=limit)
    except subprocess.TimeoutExpired:
        return None
    return re.sub(r"Time: [0-9.]+", "Time: X", r.stdout).replace(binary, "ries") + "exit=%d\n" % r.returncode

bad = skipped = 0
t0 = time.time()
for args in fuzz_cli.cases(n, seed):
    want = run(ref, args, 20)
    if want is None:
        skipped += 1
        print("SKIP (reference > 20 s):", " ".join(args), flush=True)
        continue
    got = run(exe, args, 120)
    if got != want:
        bad += 1
        print("DIFF%s:" % (" (product timed out)" if got is None else ""), " ".join(args), flush=True)
print("%d cases, %d differ, %d skipped, %.0f s" % (n, bad, skipped, time.time() - t0))
