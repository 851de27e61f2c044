"""ries_b200 -- B200-native RIES search engine behind the reference's own boundary.

The reference (thomasahle/ries) reaches its engine as ``callMain([...flags, target])`` on the wasm
build of ``public/ries.c`` and parses what it prints (src/workers/ries.worker.js:226-234).  This package
is the same call over the C ABI of ``libries_b200.so`` (include/ries_b200.h): every search stage runs in
hand-written sm_100a CUDA kernels; there is no CPU search path and importing the engine fails loudly if
the shared library is missing.
"""
from .api import (RiesError, batch, batch_shared_stats, call_main, dist_finalize, dist_init, dist_unique_id, engine_name,  # noqa: F401
                  fp64_peak_tflops, last_stats, lib_path, run)

__all__ = ["batch", "batch_shared_stats", "call_main", "run", "last_stats", "engine_name", "fp64_peak_tflops", "lib_path", "RiesError",
           "dist_unique_id", "dist_init", "dist_finalize"]
