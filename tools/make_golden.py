#!/usr/bin/env python3
"""Generates tests/golden/*.txt by running the UNMODIFIED reference (oracle/_ref/ries = gcc -O2 of
/root/reference/public/ries.c, stock glibc) on every case of tests/golden/cases.json.  The reference
cannot travel to the GPU box, so its outputs are committed as fixtures.  `Time:` is normalised.
usage: tools/make_golden.py [--big] [--huge]   (--big also regenerates the minutes-long -l5/-l6 cases,
--huge the -l7 one: 47 minutes of one host core)"""
import json, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "ries")
cases = json.load(open(os.path.join(GOLD, "cases.json")))
big = "--big" in sys.argv
env = dict(os.environ, HOME="/nonexistent")
only = [a for a in sys.argv[1:] if not a.startswith("--")]  # optional: case names to (re)generate
for c in cases:
    if only and c["name"] not in only:
        continue
    if c.get("big") and not big:
        continue
    if c.get("huge") and "--huge" not in sys.argv:
        continue
    r = subprocess.run([REF] + c["args"] + ["--no-slow-messages"], env=env, capture_output=True, text=True)
    text = re.sub(r"Time: [0-9.]+", "Time: X", r.stdout).replace(REF, "ries")
    # -Dy traces: the elapsed time of the "Still searching" line, and the average depth of the reference's
    # unbalanced tree (the product's table is a sorted array: a listed difference)
    text = re.sub(r"and [0-9.]+ sec so far", "and X sec so far", text)
    text = re.sub(r"Avg\. Depth: (?!-nan).*", "Avg. Depth: N", text)
    with open(os.path.join(GOLD, c["name"] + ".txt"), "w") as f:
        f.write(text)
        f.write("exit=%d\n" % r.returncode)
    print(c["name"], r.returncode, len(text))
