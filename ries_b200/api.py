"""ctypes binding of include/ries_b200.h (the same stub INTEGRATION.md shows for other hosts)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("RIES_B200_LIB") or os.path.join(_HERE, "libries_b200.so")
_lib = None

_WRITE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_char), ctypes.c_size_t)


class RiesError(RuntimeError):
    pass


class _Stats(ctypes.Structure):
    _fields_ = [
        ("lhs_expressions", ctypes.c_double), ("rhs_expressions", ctypes.c_double),
        ("lhs_dead_ends", ctypes.c_double), ("rhs_dead_ends", ctypes.c_double),
        ("lhs_distinct", ctypes.c_longlong), ("rhs_distinct", ctypes.c_longlong),
        ("lhs_max_complexity", ctypes.c_int), ("rhs_max_complexity", ctypes.c_int),
        ("passes", ctypes.c_int), ("exit_code", ctypes.c_int),
        ("seconds", ctypes.c_double), ("engine_seconds", ctypes.c_double), ("forms_seconds", ctypes.c_double),
        ("kernel_launches", ctypes.c_ulonglong),
        ("h2d_bytes", ctypes.c_ulonglong), ("d2h_bytes", ctypes.c_ulonglong),
        ("stage_ms", ctypes.c_double * 4), ("stage_launches", ctypes.c_ulonglong * 4),
        ("sort_keys", ctypes.c_ulonglong), ("merge_rows", ctypes.c_ulonglong),
        ("expand_nodes", ctypes.c_ulonglong),
        ("op_char", ctypes.c_char * 48), ("op_count_rhs", ctypes.c_ulonglong * 48),
        ("op_count_lhs", ctypes.c_ulonglong * 48),
        ("sharded_passes", ctypes.c_ulonglong), ("gather_bytes", ctypes.c_ulonglong),
        ("gather_ms", ctypes.c_double),
    ]


def lib_path():
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RiesError(
                "ries_b200: %s is missing -- build it with `python -m ries_b200.build` "
                "(there is no CPU fallback)" % _LIB_PATH)
        lib = ctypes.CDLL(_LIB_PATH)
        lib.ries_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), _WRITE_FN, ctypes.c_void_p]
        lib.ries_main.restype = ctypes.c_int
        lib.ries_last_stats.argtypes = [ctypes.POINTER(_Stats)]
        lib.ries_last_stats.restype = None
        lib.ries_fp64_peak_tflops.argtypes = []
        lib.ries_fp64_peak_tflops.restype = ctypes.c_double
        lib.ries_dist_unique_id.argtypes = [ctypes.c_char_p]
        lib.ries_dist_unique_id.restype = ctypes.c_int
        lib.ries_dist_init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        lib.ries_dist_init.restype = ctypes.c_int
        lib.ries_dist_finalize.argtypes = []
        lib.ries_dist_finalize.restype = None
        lib.ries_engine_name.argtypes = []
        lib.ries_engine_name.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def call_main(args, prog="ries"):
    """``callMain(args)`` of the wasm module: returns (exit_code, everything printed)."""
    lib = _load()
    argv = [prog.encode()] + [a.encode() if isinstance(a, str) else a for a in args]
    arr = (ctypes.c_char_p * (len(argv) + 1))(*argv, None)
    chunks = []

    def _write(_ctx, buf, n):
        chunks.append(ctypes.string_at(buf, n))

    cb = _WRITE_FN(_write)
    code = lib.ries_main(len(argv), arr, cb, None)
    return code, b"".join(chunks).decode("utf-8", "replace")


def run(target, *flags):
    """Convenience: ``run("2.5063", "-l4", "-F3")`` -> printed report (raises on usage errors)."""
    code, text = call_main(list(flags) + [str(target)])
    if code not in (0, -9998):
        raise RiesError("ries exited with status %d:\n%s" % (code, text))
    return text


def last_stats():
    s = _Stats()
    _load().ries_last_stats(ctypes.byref(s))
    out = {}
    for name, _ in _Stats._fields_:
        v = getattr(s, name)
        if name == "op_char":
            out[name] = bytes(s.op_char)
        else:
            out[name] = list(v) if hasattr(v, "__len__") else v
    chars = out.pop("op_char")
    out["exec_calls"] = {chr(c): {"rhs": out["op_count_rhs"][i], "lhs": out["op_count_lhs"][i]}
                         for i, c in enumerate(chars) if c and (out["op_count_rhs"][i] or out["op_count_lhs"][i])}
    del out["op_count_rhs"], out["op_count_lhs"]
    return out


_BATCH_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char),
                             ctypes.c_size_t)


def batch(targets, flags=("-F3",), workers=8):
    """BASELINE config 5: many independent targets with the same flags.  Returns [(exit_code, text), ...] in
    the order of `targets`.  `workers` host threads drive one device engine + CUDA stream each."""
    lib = _load()
    lib.ries_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                               ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, _BATCH_FN, ctypes.c_void_p]
    lib.ries_batch.restype = ctypes.c_int
    t_arr = (ctypes.c_char_p * len(targets))(*[str(t).encode() for t in targets])
    f_arr = (ctypes.c_char_p * max(len(flags), 1))(*[f.encode() for f in flags])
    out = [None] * len(targets)

    def _done(_ctx, index, code, buf, n):
        out[index] = (code, ctypes.string_at(buf, n).decode("utf-8", "replace"))

    cb = _BATCH_FN(_done)
    lib.ries_batch(len(targets), t_arr, len(flags), f_arr, workers, cb, None)
    return out


def batch_shared_stats():
    """RHS passes enumerated once for the last batch, and RHS passes of its targets served from them."""
    lib = _load()
    a, b = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    lib.ries_batch_shared_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]
    lib.ries_batch_shared_stats.restype = None
    lib.ries_batch_shared_stats(ctypes.byref(a), ctypes.byref(b))
    return {"rhs_passes_enumerated_once": a.value, "rhs_passes_served": b.value}


def fp64_peak_tflops():
    return _load().ries_fp64_peak_tflops()


def engine_name():
    return _load().ries_engine_name().decode()


def dist_unique_id():
    """Rank 0: a fresh 128-byte NCCL id for `dist_init` (hand it to the other ranks over any side channel)."""
    buf = ctypes.create_string_buffer(128)
    rc = _load().ries_dist_unique_id(buf)
    if rc != 0:
        raise RiesError("ries_dist_unique_id failed (%d)" % rc)
    return buf.raw


def dist_init(rank, world, unique_id):
    """Every rank (one process per GPU, RIES_DEVICE already set): join the communicator.  Afterwards every
    rank calls `call_main` with the same arguments and they search together (large passes are sharded)."""
    rc = _load().ries_dist_init(int(rank), int(world), bytes(unique_id))
    if rc != 0:
        raise RiesError("ries_dist_init failed (%d)" % rc)


def dist_finalize():
    _load().ries_dist_finalize()
