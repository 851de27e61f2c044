import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def normalise(text, prog=None):
    text = re.sub(r"Time: [0-9.]+", "Time: X", text)
    # -Dy traces (ries.c:12512-12530): elapsed seconds, and bt_stats' average tree depth (listed difference)
    text = text.replace("\r", "\n")  # the goldens were captured in text mode (universal newlines)
    text = re.sub(r"and [0-9.]+ sec so far", "and X sec so far", text)
    text = re.sub(r"Avg\. Depth: (?!-nan).*", "Avg. Depth: N", text)
    if prog:
        text = text.replace(prog, "ries")
    return text


def golden_cases(big=None):
    cases = json.load(open(os.path.join(GOLD, "cases.json")))
    out = []
    for c in cases:
        if big is not None and bool(c.get("big")) != big:
            continue
        if os.path.exists(os.path.join(GOLD, c["name"] + ".txt")):
            out.append(c)
    return out


def golden_text(name):
    with open(os.path.join(GOLD, name + ".txt")) as f:
        return f.read()


def run_binary(path, args, timeout=900):
    env = dict(os.environ, HOME="/nonexistent")
    r = subprocess.run([path] + list(args) + ["--no-slow-messages"], env=env, capture_output=True, text=True,
                       timeout=timeout)
    return normalise(r.stdout, path) + "exit=%d\n" % r.returncode


@pytest.fixture(scope="session")
def oracle_build():
    """Builds the checker binaries (C restatement, CPU emulation, rmath twin; reference binaries only
    where /root/reference exists -- on the GPU box the prebuilt ones travel with the snapshot)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    return os.path.join(ROOT, "oracle")


@pytest.fixture(scope="session")
def product_lib():
    from ries_b200 import build
    return build.build()
