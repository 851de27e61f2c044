"""CPU-only: the measurement code of bench.py that does not need a GPU -- the CPU arm's helpers, the batch target
stream, the golden look-ups the B200 arm checks its reports against."""
import json
import os
import random
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_batch_target_stream_is_the_survey_stream():
    rng = random.Random(20261019)
    want = ["%.6f" % rng.uniform(0, 10) for _ in range(8)]
    assert bench.targets_for(8) == want
    with open(os.path.join(ROOT, "tests", "golden", "batch_l4.json")) as f:
        g = json.load(f)
    assert bench.targets_for(64) == g["targets"] and len(g["reports"]) == 64


def test_side_configs_have_goldens():
    for _key, gname, argv in bench.SIDE_CONFIGS:
        assert bench.golden_text(gname) is not None, gname
        assert (argv or bench.golden_args(gname)) is not None, gname
    assert bench.golden_text("l6full") is not None


@pytest.mark.skipif(bench.ref_binary()[0] is None, reason="no CPU binary built")
def test_cpu_arm_helpers(tmp_path, monkeypatch):
    binary, kind = bench.ref_binary()
    argv = ["2.5063", "-l1", "-F3"]
    # the full run, and the per-box cache the other invocations of a session read
    monkeypatch.setattr(bench, "cpu_cache_path", lambda b, a: str(tmp_path / "cache.json"))
    c = bench.cpu_full_run(binary, kind, argv)
    assert c["expressions"] == 130334 and c["seconds"] > 0 and not c["cached"]
    c2 = bench.cpu_full_run(binary, kind, argv)
    assert c2["cached"] and c2["expressions"] == c["expressions"] and c2["seconds"] == c["seconds"]
    # the bounded sample: expressions from the reference's own per-pass trace
    p = bench.cpu_bounded_prefix(binary, ["2.5063", "-l2", "-F3"], 20.0)
    assert p["ran_to_completion"] and p["expressions"] >= 512000 and p["seconds"] > 0


def test_report_normalisation_and_status():
    assert bench.normalise("x  Time: 0.123\n") == "x  Time: X\n"
    assert bench.shown_status(-9998) == 242 and bench.shown_status(0) == 0
    assert bench.parse_expressions("        expressions:       55283        75051       130334\n") == 130334
