"""CPU-only: the checkers against the reference's own outputs (tests/golden, made by tools/make_golden.py
from the unmodified ries.c) and against each other."""
import os
import subprocess

import pytest

from conftest import ROOT, golden_cases, golden_text, run_binary

SMALL = [c for c in golden_cases(big=False) if c["name"] not in ("t25063_l4_F3", "infix_l4")]
# the C restatement covers the postfix/FORTH formats (-F0/-F3); infix is checked through the emulation + GPU
# ... and not --no-refinement (host-only bookkeeping shared by the emulation and the product)
POSTFIX = [c for c in SMALL if not c["name"].startswith(("infix_", "cfg1_", "solve_", "web_", "norefine_", "opt_"))
           and "infix" not in c["name"]]


@pytest.mark.parametrize("case", SMALL, ids=[c["name"] for c in SMALL])
def test_batch_emulation_matches_reference_goldens(oracle_build, case):
    """The batch formulation (generate pass -> stable sort -> dedupe -> merge -> gap-local neighbour test ->
    ordered replay) driven on the CPU with rmath must print exactly what the streaming reference prints."""
    got = run_binary(os.path.join(oracle_build, "_build", "ries_emul"), case["args"])
    assert got == golden_text(case["name"])


@pytest.mark.parametrize("case", POSTFIX, ids=[c["name"] for c in POSTFIX])
def test_c_restatement_matches_reference_goldens(oracle_build, case):
    """oracle/ries_oracle.c (streaming restatement, stock libm) is pinned to the reference's outputs."""
    got = run_binary(os.path.join(oracle_build, "_build", "ries_oracle"), case["args"])
    assert got == golden_text(case["name"])


@pytest.mark.parametrize("name", ["t25063_l3_F3_x", "i143_l3", "neg_l2"])
def test_pass_fed_in_slices_prints_the_same(oracle_build, name):
    """A pass with more forms than the engine's 16-bit form index is fed in consecutive slices (extreme
    --symbol-weights); with the test hook every pass is sliced, and nothing may change."""
    case = [c for c in golden_cases() if c["name"] == name][0]
    env = dict(os.environ, HOME="/nonexistent", RIES_FORMS_SLICE="5")
    exe = os.path.join(oracle_build, "_build", "ries_emul")
    r = subprocess.run([exe] + case["args"] + ["--no-slow-messages"], env=env, capture_output=True, text=True)
    from conftest import normalise
    assert normalise(r.stdout, exe) + "exit=%d\n" % r.returncode == golden_text(name)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ries")), reason="no reference build")
def test_random_command_lines_emulation_vs_reference(oracle_build):
    """Differential fuzz of the host logic (argument parsing, set-up, replay, formatting) shared by the product:
    seeded random option combinations through the emulation and through the unmodified reference."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_cli
    ref = os.path.join(oracle_build, "_ref", "ries")
    emu = os.path.join(oracle_build, "_build", "ries_emul")
    for args in fuzz_cli.cases(25, 42):
        assert run_binary(emu, args) == run_binary(ref, args), " ".join(args)


def test_l4_headline_config(oracle_build):
    """BASELINE config 2 (ries 2.5063 -l4): emulation with rmath == unmodified reference, all 22 equations and
    every counter of the stats block."""
    case = [c for c in golden_cases() if c["name"] == "t25063_l4_F3"][0]
    got = run_binary(os.path.join(oracle_build, "_build", "ries_emul"), case["args"])
    assert got == golden_text(case["name"])
    assert "8948692" in got and "127357792" in got and "2635817" in got  # SURVEY section 8(c) KATs


@pytest.mark.parametrize("func", ["exp", "log", "pow", "sincos", "tan", "atan2"])
def test_rmath_bit_exact_vs_platform_libm(oracle_build, func):
    """rmath must reproduce glibc's libm bit for bit on the ranges exec() can reach (and well beyond)."""
    r = subprocess.run([os.path.join(oracle_build, "_build", "rmath_diff"), func, "3000000", "11"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert ": 0 / 3000000 mismatches" in r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ries_rm")), reason="no reference build")
@pytest.mark.parametrize("args", [["2.5063", "-l2", "-F3"], ["0.5772156649", "-l2", "-F0"], ["19.99", "-l2", "-F3"]])
def test_reference_with_rmath_equals_reference_with_glibc(oracle_build, args):
    """The unmodified reference source linked against rmath prints what it prints with glibc."""
    a = run_binary(os.path.join(oracle_build, "_ref", "ries"), args)
    b = run_binary(os.path.join(oracle_build, "_ref", "ries_rm"), args)
    assert a == b
