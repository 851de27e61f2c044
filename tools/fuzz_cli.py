#!/usr/bin/env python3
"""Differential fuzz of the command line: random combinations of the supported options and targets through the
unmodified reference (oracle/_ref/ries) and through the CPU emulation (oracle/_build/ries_emul, which shares
every line of host code and per-thread math with the product).  Prints the cases that differ.
usage: tools/fuzz_cli.py [n_cases] [seed]"""
import os, random, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = os.path.join(ROOT, "oracle", "_ref", "ries")
EMU = os.path.join(ROOT, "oracle", "_build", "ries_emul")
rng = random.Random(1)
SYMS = "123456789pefnrsqlESCT+-*/^vLA"

def target():
    k = rng.random()
    if k < 0.5: return "%.*f" % (rng.randint(2, 9), rng.uniform(0.05, 30))
    if k < 0.6: return str(rng.randint(2, 200))
    if k < 0.7: return "-%.5f" % rng.uniform(0.1, 9)
    if k < 0.8: return "%.4e" % rng.uniform(1e-4, 1e5)
    return rng.choice(["3.14159265358979", "1.6180339887", "2.718281828", "0.5772156649", "137.035999", "1.4142135"])

def flags():
    f = []
    f.append("-l%s" % rng.choice(os.environ.get("FUZZ_LEVELS", "0 1 1 2 0.5 -1").split()))
    if rng.random() < 0.7: f.append(rng.choice(["-F0", "-F1", "-F2", "-F3", "-F"]))
    if rng.random() < 0.25: f.append(rng.choice(["-i", "-r", "-a", "-c", "-ie", "-re", "-l"]))
    if rng.random() < 0.25: f.append("-N" + "".join(rng.sample(SYMS, rng.randint(1, 6))))
    if rng.random() < 0.15: f.append("-O" + "".join(rng.sample(SYMS, rng.randint(1, 5))))
    if rng.random() < 0.10: f.append("-S" + "".join(rng.sample(SYMS, rng.randint(6, 16))))
    if rng.random() < 0.10: f.append("-E" + "".join(rng.sample(SYMS, rng.randint(1, 5))))
    if rng.random() < 0.2: f.append("-s")
    if rng.random() < 0.15: f.append("-x")
    if rng.random() < 0.1: f.append("--wide")
    if rng.random() < 0.1: f += ["-n%d" % rng.randint(1, 30)]
    if rng.random() < 0.1: f.append("--one-sided")
    if rng.random() < 0.1: f.append("--no-refinement")
    if rng.random() < 0.1: f.append("--mad")
    if rng.random() < 0.1: f.append("--explicit-multiply")
    if rng.random() < 0.1: f += ["--max-match-distance", rng.choice(["1e-3", "-0.001", "0", "0.5"])]
    if rng.random() < 0.08: f += ["--min-match-distance", rng.choice(["1e-6", "-1e-5", "0"])]
    if rng.random() < 0.08: f += ["--max-trig-cycles", rng.choice(["2", "0.5", "10"])]
    if rng.random() < 0.08: f += ["--significance-loss-margin", rng.choice(["1", "3", "6"])]
    if rng.random() < 0.08: f += ["--derivative-margin", rng.choice(["1e-7", "1e-5", "1e-9"])]
    if rng.random() < 0.08: f += ["--max-equate-value", rng.choice(["10", "100"])]
    if rng.random() < 0.05: f += ["--numeric-anagram", "".join(rng.choice("123456789") for _ in range(rng.randint(2, 5)))]
    if rng.random() < 0.06: f += ["--trig-argument-scale", rng.choice(["1", "0.0174532925199433", "2", "3.141592653589793", "0"])]
    if rng.random() < 0.04: f += ["--symbol-weights", rng.choice(["8:T", "20:^", "14:x", "5:p"])]
    if rng.random() < 0.05: f.append("--no-canon-simplify")
    if rng.random() < 0.12: f += ["--canon-reduction", rng.choice(["nr25", "n", "r", "2", "5", "nr", "25", "r5", "n2", "0"])]
    if rng.random() < 0.05: f += ["--any-exponents"]
    if rng.random() < 0.05: f += ["--rational-exponents"]
    rng.shuffle(f)
    return f

def run(exe, args):
    try:
        r = subprocess.run([exe] + args + ["--no-slow-messages"], capture_output=True, text=True,
                           env=dict(os.environ, HOME="/nonexistent"), timeout=120)
    except subprocess.TimeoutExpired:
        return "TIMEOUT"
    t = re.sub(r"Time: [0-9.]+", "Time: X", r.stdout).replace(exe, "ries")
    return t + "exit=%d\n" % r.returncode

def cases(n, seed):
    """n random argv lists (deterministic for a seed)."""
    global rng
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        t = target()
        args = flags()
        args.insert(rng.randint(0, len(args)), t)
        # options added later draw from a stream of their own, so that the case lists of the seeds the test suites
        # pin (tests/test_oracle_cpu.py, tests/test_gpu_parity.py) stay what they were apart from the additions
        r2 = random.Random(" ".join(args))
        if r2.random() < 0.03:
            k = r2.randint(0, len(args)); args[k:k] = ["--min-memory", r2.choice(["1e6", "0", "-3e9"])]
        if r2.random() < 0.03:
            k = r2.randint(0, len(args)); args[k:k] = ["--memory-abort-threshold", r2.choice(["1.5", "3", "1.0", "x"])]
        out.append(args)
    return out


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
  seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
  bad = 0
  for args in cases(n, seed):
    a, b = run(REF, args), run(EMU, args)
    if a != b:
        bad += 1
        print("DIFF:", " ".join(args))
        la, lb = a.splitlines(), b.splitlines()
        for x, y in zip(la, lb):
            if x != y:
                print("   ref:", x[:150]); print("   emu:", y[:150]); break
        else:
            print("   line counts", len(la), len(lb))
  print("%d cases, %d differ" % (n, bad))


if __name__ == "__main__":
    main()
