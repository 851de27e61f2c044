#!/usr/bin/env python3
"""Differential fuzz on a GPU box: random command lines (tools/fuzz_cli.py) through the product (C ABI) and
through the prebuilt unmodified reference oracle/_ref/ries.  usage: tools/gpu_fuzz.py [n] [seed]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_cli, ries_b200
from conftest import normalise, run_binary
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ref = os.path.join(ROOT, "oracle", "_ref", "ries")
bad = 0
for args in fuzz_cli.cases(n, seed):
    code, text = ries_b200.call_main(args + ["--no-slow-messages"])
    got = normalise(text, "ries") + "exit=%d\n" % (code & 0xff if code != 0 else 0)
    want = run_binary(ref, args)
    if got != want:
        bad += 1
        print("DIFF:", " ".join(args))
print("%d cases, %d differ" % (n, bad))
