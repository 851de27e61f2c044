"""CPU-only: the C-ABI library loads, exports what include/ries_b200.h declares, and refuses to search
without a GPU (there is no CPU path in the product)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ries_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ries_[a-z0-9_]+)\s*\(", text)) - {"ries_write_fn", "ries_batch_done_fn"})


def test_header_declares_the_boundary():
    assert declared_functions() == ["ries_batch", "ries_batch_shared_stats", "ries_dist_finalize", "ries_dist_init", "ries_dist_unique_id",
                                    "ries_engine_name", "ries_fp64_peak_tflops", "ries_last_stats", "ries_main"]


def test_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(product_lib)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_usage_errors_do_not_need_a_gpu(product_lib):
    import ries_b200
    code, text = ries_b200.call_main([])
    assert code == 1 and "Please specify a target number." in text
    code, text = ries_b200.call_main(["2.5", "--bogus"])
    assert code == -1 and "Unknown option '--bogus'" in text


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ries_b200
    assert ries_b200.engine_name().startswith("unavailable")
    code, text = ries_b200.call_main(["2.5063", "-l0", "-F3"])
    assert code == 3
    assert " = " not in text  # no equation was produced


def test_product_does_not_reference_the_oracle():
    """Nothing under ries_b200/ may include, link or execute oracle/ (oracle -> product is the only direction)."""
    for root, _, files in os.walk(os.path.join(ROOT, "ries_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".c")):
                text = open(os.path.join(root, f), errors="replace").read()
                for line in text.splitlines():
                    if "oracle/" in line and not line.lstrip().startswith(("//", "*", "/*", "#", '"""')):
                        stripped = line.strip()
                        assert stripped.startswith(("//", "/*", "*")) or "oracle/" in stripped.split("//")[-1], (f, line)
