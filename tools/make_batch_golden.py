#!/usr/bin/env python3
"""Generates tests/golden/batch_l4.json: the UNMODIFIED reference (oracle/_ref/ries) run on the first 64
targets of BASELINE config 5's stream (Python random.seed(20261019), "%.6f" % uniform(0,10), zero rejected)
with `-l4 -F3`, one process per target (the reference has no batch mode).
`Time:` is normalised; the exit status is appended as in the other goldens.
usage: tools/make_batch_golden.py [n_targets] [level]"""
import json, os, random, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = os.path.join(ROOT, "oracle", "_ref", "ries")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
level = sys.argv[2] if len(sys.argv) > 2 else "4"
rng = random.Random(20261019)
targets = []
while len(targets) < n:
    t = "%.6f" % rng.uniform(0, 10)
    if float(t) != 0.0:
        targets.append(t)
env = dict(os.environ, HOME="/nonexistent")


def one(t):
    r = subprocess.run([REF, t, "-l" + level, "-F3", "--no-slow-messages"], env=env, capture_output=True, text=True)
    return re.sub(r"Time: [0-9.]+", "Time: X", r.stdout).replace(REF, "ries") + "exit=%d\n" % r.returncode


with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
    texts = list(ex.map(one, targets))
out = {"seed": 20261019, "flags": ["-l" + level, "-F3"], "targets": targets, "reports": texts}
path = os.path.join(ROOT, "tests", "golden", "batch_l%s.json" % level)
with open(path, "w") as f:
    json.dump(out, f, indent=0)
print(path, len(targets), sum(len(t) for t in texts))
