#!/usr/bin/env python3
"""bench.py -- expressions evaluated/sec and time-to-full-report of the RIES search on B200.

A "step" is one complete search (enumerate -> evaluate -> sort/dedupe/merge -> match/Newton -> report)
through the reference-facing C ABI (`ries_main(argc, argv, write_cb)`, include/ries_b200.h), i.e. exactly
the call the web worker makes (`callMain([...flags, target])`).  Nothing is reused between steps except
device allocations: every step rebuilds the whole value table from scratch.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N independent targets)
    python bench.py --impl reference          (the reference's own CPU implementation, oracle/_ref/ries)

Default workload (N=1): `ries 2.5063141592653589 -l6 -F3`, the full-depth -l6 search BASELINE.json's metric
is quoted on (1.45e8 expressions, 2.04e9 dead-ends, 3.23e7 table rows; fits one GPU).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_TARGET = "2.5063141592653589"
DEFAULT_LEVEL = "6"
# BASELINE.md config 5 distribution: f"{random.uniform(0,10):.6f}", seed 20261019 (first 8 targets)
BATCH_SEED = 20261019

# algorithmic bytes per survivor through the sort (read once + write once of a
# 32-byte row) and per merged row (read old+new, write merged); flop weights per exec'd symbol.
ROW_BYTES = 32
NODE_BYTES = 48
SORT_BYTES_PER_KEY = 2 * ROW_BYTES
MERGE_BYTES_PER_ROW = 2 * ROW_BYTES
# fp64 flops per exec() of each symbol: (value, extra for d/dx), SURVEY.md section 8(d)
FLOP_WEIGHTS = {"n": (1, 1), "s": (1, 2), "r": (1, 3), "q": (1, 2), "+": (7, 1), "-": (7, 1), "*": (1, 3), "/": (1, 5),
                "l": (25, 1), "E": (20, 1), "S": (50, 3), "C": (50, 3), "T": (45, 4), "^": (60, 30), "v": (62, 32),
                "L": (52, 8), "A": (45, 7)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def ref_binary():
    for name in ("ries",):
        p = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(p):
            return p, "reference"
    p = os.path.join(ROOT, "oracle", "_build", "ries_oracle")
    return (p, "port") if os.path.exists(p) else (None, None)


def parse_expressions(text):
    m = re.search(r"expressions:\s+(\d+)\s+(\d+)\s+(\d+)", text)
    return int(m.group(3)) if m else 0


def run_cpu(binary, argv):
    env = dict(os.environ, HOME="/nonexistent")
    t0 = time.perf_counter()
    r = subprocess.run([binary] + argv + ["--no-slow-messages"], env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    return parse_expressions(r.stdout), dt


def targets_for(n):
    import random
    rng = random.Random(BATCH_SEED)
    out = []
    while len(out) < n:
        t = "%.6f" % rng.uniform(0, 10)
        if float(t) != 0.0:
            out.append(t)
    return out


def batch_workload(args, rank, local_rank, world, dist, torch, ries_b200):
    """BASELINE config 5: targets ~ U(0,10) with 6 decimals (seed 20261019) at -l<level>, sharded round-robin
    over ranks, `--workers` concurrent engines per GPU.  A step = the whole batch once."""
    from ries_b200.parallel import shard_targets
    level = args.level if args.level != DEFAULT_LEVEL else "4"
    mine = shard_targets(targets_for(args.batch_size), rank, world)
    flags = ["-l" + level, "-F3", "--no-slow-messages"]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        ries_b200.batch(mine[:max(2 * args.workers, 8)], flags, workers=args.workers)
    barrier()
    t0 = time.perf_counter()
    exprs = 0
    for _ in range(args.steps):
        for code, text in ries_b200.batch(mine, flags, workers=args.workers):
            exprs += parse_expressions(text)
    barrier()
    wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([float(exprs), float(len(mine) * args.steps)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        tm = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        exprs_all, searches_all = t.tolist()
        wall_all = tm.item()
    else:
        exprs_all, searches_all, wall_all = float(exprs), float(len(mine) * args.steps), wall
    if rank == 0:
        print(json.dumps({
            "metric": "expressions_evaluated_per_sec", "value": exprs_all / wall_all, "unit": "expr/s",
            "n_gpus": max(args.gpus, world), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batch of %d targets U(0,10) 6dp seed %d, ries <t> -l%s -F3, %d engines per GPU"
                                   % (args.batch_size, BATCH_SEED, level, args.workers)},
            "targets_per_sec": searches_all / wall_all,
            "e2e": {"value": exprs_all / wall_all, "unit": "expr/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "api": "ries_batch(targets, flags, workers, done_cb) via ctypes"},
        }))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--target", default=DEFAULT_TARGET)
    ap.add_argument("--level", default=DEFAULT_LEVEL)
    ap.add_argument("--cpu-level", default="4", help="level of the bounded CPU sample (same target)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="search", choices=["search", "batch"],
                    help="search: one full-depth search per step (default); batch: BASELINE config 5, many targets")
    ap.add_argument("--batch-size", type=int, default=256, help="targets per step over all ranks (batch workload)")
    ap.add_argument("--workers", type=int, default=8, help="concurrent engines per GPU (batch workload)")
    ap.add_argument("--shard-search", action="store_true",
                    help="N>1: all ranks work on ONE search (K1 sharded by forms, NCCL all-gather of the complete "
                         "expressions; strong scaling) instead of one independent target per rank")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    workload = "ries %s -l%s -F3" % (args.target, args.level)
    sample_desc = ("the first pass-prefix of the same search: `ries %s -l%s` (identical pass sequence, stops at "
                   "the -l%s expression budget)" % (args.target, args.cpu_level, args.cpu_level))

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        binary, kind = ref_binary()
        if binary is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ries and oracle/_build/ries_oracle are both missing"}))
            return 0
        argv = [args.target, "-l" + args.cpu_level, "-F3"]
        for _ in range(min(args.warmup, 1)):
            run_cpu(binary, argv)
        total_expr, total_t = 0, 0.0
        for _ in range(args.steps):
            n, dt = run_cpu(binary, argv)
            total_expr += n
            total_t += dt
        v = total_expr / total_t
        line = {
            "impl": "reference", "metric": "expressions_evaluated_per_sec", "value": v, "unit": "expr/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample_desc},
            "cpu_baseline": {"value": v, "unit": "expr/s", "cores": 1, "kind": kind, "sample": sample_desc},
            "e2e": {"value": v, "unit": "expr/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["RIES_DEVICE"] = str(local_rank)
    import ries_b200

    if args.workload == "batch":
        return batch_workload(args, rank, local_rank, world, dist, torch, ries_b200)

    def join_search_communicator():
        """ries_dist_init on every rank: the id comes from rank 0 over torch.distributed (plumbing only)."""
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ries_b200.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ries_b200.dist_init(rank, world, uid.cpu().numpy().tobytes())

    sharded = args.shard_search and world > 1
    if sharded:
        join_search_communicator()
        target = args.target
    elif world > 1:
        target = targets_for(world)[rank]  # independent targets: BASELINE config-5 style sharding
    else:
        target = args.target
    argv = [target, "-l" + args.level, "-F3", "--no-slow-messages"]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        t0 = time.perf_counter()
        code, text = ries_b200.call_main(argv)
        dt = time.perf_counter() - t0
        st = ries_b200.last_stats()
        assert code in (0, -9998), (code, text[-400:])
        return dt, st, text

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_begin = time.perf_counter()
    per_step = []
    for _ in range(args.steps):
        per_step.append(step())
    barrier()
    wall = time.perf_counter() - t_begin
    clocks = sampler.stop()

    # Cross-check of the per-stage device times: the engine takes them from device clock stamps (%globaltimer)
    # that the first kernel of every stage leaves in the pass's control block; here a few extra steps, outside
    # the timed region, bracket every stage with CUDA events on the engine's stream instead (RIES_STAGE_EVENTS).
    event_check = None
    if world == 1 and not sharded:
        try:
            os.environ["RIES_STAGE_EVENTS"] = "1"
            ev_steps = [step() for _ in range(3)]
            ev_ms = [sum(s[1]["stage_ms"][i] for s in ev_steps) / len(ev_steps) for i in range(4)]
            event_check = {"steps": len(ev_steps), "ms_per_step_wall": 1e3 * sum(s[0] for s in ev_steps) / len(ev_steps),
                           "stage_ms_per_step": {"K1_enum_eval": ev_ms[0], "K2_sort": ev_ms[1],
                                                 "K2K3_dedupe_match_merge": ev_ms[2], "K4_refine": ev_ms[3]},
                           "how": "cudaEventRecord before/after every stage on the engine's stream, cudaEventElapsedTime; "
                                  "tiny passes go through the general kernels in this mode so that the stages stay separable"}
        finally:
            os.environ.pop("RIES_STAGE_EVENTS", None)

    exprs = sum(s[1]["lhs_expressions"] + s[1]["rhs_expressions"] for s in per_step)
    dead = sum(s[1]["lhs_dead_ends"] + s[1]["rhs_dead_ends"] for s in per_step)
    dev_ms = [sum(s[1]["stage_ms"][i] for s in per_step) for i in range(4)]
    dev_launches = [sum(s[1]["stage_launches"][i] for s in per_step) for i in range(4)]
    launches = sum(s[1]["kernel_launches"] for s in per_step)
    h2d = sum(s[1]["h2d_bytes"] for s in per_step) / args.steps
    d2h = sum(s[1]["d2h_bytes"] for s in per_step) / args.steps
    sort_keys = sum(s[1]["sort_keys"] for s in per_step)
    merge_rows = sum(s[1]["merge_rows"] for s in per_step)
    nodes = sum(s[1]["expand_nodes"] for s in per_step)
    dev_total_s = sum(dev_ms) / 1e3

    # N>1, default mode: after the independent-target steps, the same ranks also time ONE search together
    # (north_star's "time-to-full-report at -l6 on N GPUs"); reported beside the weak-scaling value
    sharded_extra = None
    if dist is not None and not sharded and os.environ.get("RIES_BENCH_NO_SHARDED_EXTRA") is None:
        join_search_communicator()
        sargv = [args.target, "-l" + args.level, "-F3", "--no-slow-messages"]
        for _ in range(2):
            ries_b200.call_main(sargv)
        barrier()
        ts0 = time.perf_counter()
        for _ in range(args.steps):
            scode, _stext = ries_b200.call_main(sargv)
        barrier()
        swall = time.perf_counter() - ts0
        sst = ries_b200.last_stats()
        ries_b200.dist_finalize()
        tm = torch.tensor([swall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        sexpr = sst["lhs_expressions"] + sst["rhs_expressions"]
        sharded_extra = {
            "workload": "ries %s -l%s -F3, ONE search over %d GPUs" % (args.target, args.level, world),
            "time_to_full_report_s": tm.item() / args.steps, "expr_per_s_e2e": sexpr * args.steps / tm.item(),
            "sharded_passes": sst["sharded_passes"], "allgather_MB_per_search": sst["gather_bytes"] / 1e6, "allgather_ms": sst["gather_ms"],
            "stage_ms": sst["stage_ms"], "scaling": "strong",
            "how": "forms dealt round-robin to ranks for K1; each rank sorts and pre-dedupes its own expressions "
                   "against its table replica; the kept rows are all-gathered rank-major over NCCL; the final "
                   "sort/dedupe/match/merge run replicated; every rank prints the identical report",
        }

    # whole-job numbers: sum of units over ranks, max of time over ranks
    if dist is not None:
        if sharded:  # every rank holds the totals of the one search: count them once
            exprs, dead = (exprs, dead) if rank == 0 else (0.0, 0.0)
        t = torch.tensor([exprs, dead, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        tmax = torch.tensor([wall, dev_total_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        exprs_all, dead_all, launches_all = t.tolist()
        wall_all, dev_all = tmax.tolist()
    else:
        exprs_all, dead_all, launches_all, wall_all, dev_all = exprs, dead, launches, wall, dev_total_s

    if sharded:
        ries_b200.dist_finalize()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_kind = peaks()
    # ---- roofline of the dominant stage (device time from CUDA events on the engine's stream) ----
    # algorithmic bytes: K1 streams every live prefix once out (48 B) and once in (48 B)
    # and writes each complete expression once (32 B row + 2 B form); the sort reads and writes each 32 B row
    # once; the merge reads old+new and writes merged (2 x 32 B per merged row) + one 32 B neighbour read.
    k1_bytes = nodes * 2 * NODE_BYTES + sort_keys * (ROW_BYTES + 2)
    sort_bytes = sort_keys * SORT_BYTES_PER_KEY
    merge_bytes = merge_rows * MERGE_BYTES_PER_ROW + merge_rows * ROW_BYTES
    stages = [
        ("K1 enumerate+evaluate (k_expand)", k1_bytes, dev_ms[0], dev_launches[0]),
        ("K2 order+radix sort (k_reorder/k_rs_onesweep)", sort_bytes, dev_ms[1], dev_launches[1]),
        ("K2/K3 classify+match+merge (k_classify_compact/k_match_merge_new/k_merge_old)", merge_bytes, dev_ms[2], dev_launches[2]),
    ]
    dom = max(stages, key=lambda s: s[2])
    achieved = dom[1] / (dom[2] / 1e3) / 1e9 if dom[2] > 0 else 0.0
    # K1's fp64 figure: exec() calls per symbol (counted on the device) x SURVEY 8(d) flop weights
    flops = 0.0
    execs = 0
    for st in (s[1] for s in per_step):
        for ch, c in st["exec_calls"].items():
            w, wdx = FLOP_WEIGHTS.get(ch, (0, 0))
            flops += c["rhs"] * w + c["lhs"] * (w + wdx)
            execs += c["rhs"] + c["lhs"]
    fp64_peak = ries_b200.fp64_peak_tflops()
    fp64_ach = flops / (dev_ms[0] / 1e3) / 1e12 if dev_ms[0] > 0 else 0.0
    # DRAM traffic of the dominant kernel: not measurable outside a profiler, so it is taken from the committed
    # ncu capture of this same workload (tools/gpu_profile.sh -> profiles/), per launch like `achieved`
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_l6_k_expand_traffic.json")
    if dom[0].startswith("K1") and args.workload == "search" and args.target == DEFAULT_TARGET \
            and args.level == DEFAULT_LEVEL and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_launch"]
        traffic_src = "profiles/r1_l6_k_expand_traffic.json: ncu dram__bytes_read.sum+dram__bytes_write.sum over the %d K1 launches (k_expand + k_tiny_front) of one search (cold process), per launch" % tj["launches"]
    roofline = {
        "bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
        "algorithmic_bytes_per_launch": (dom[1] / dom[3]) if dom[3] else None,
        "algorithmic_bytes_per_step": dom[1] / args.steps, "launches_per_step": dom[3] / args.steps,
        "avg_launch_ms": dom[2] / dom[3] if dom[3] else None,
        "stage_ms_per_step": {"K1_enum_eval": dev_ms[0] / args.steps, "K2_sort": dev_ms[1] / args.steps,
                              "K2K3_dedupe_match_merge": dev_ms[2] / args.steps, "K4_refine": dev_ms[3] / args.steps},
        "stage_hbm_frac": {s[0].split(" ")[0] + "_" + str(i): (s[1] / (s[2] / 1e3) / 1e9 / hbm_peak if s[2] > 0 else 0.0)
                           for i, s in enumerate(stages)},
        "k1_fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": fp64_ach / fp64_peak if fp64_peak else None, "peak_kind": "measured DFMA microbenchmark",
                    "exec_calls_per_step": execs / args.steps, "algorithmic_flop_per_step": flops / args.steps},
        "stage_timing": "device clock stamps (%globaltimer) written by the first kernel of each stage into the pass's control block, on the engine's stream, summed over the timed region",
        "cuda_event_check": event_check,
        "note": "K1 does ~14 flop per exec() and ~10 integer tests per exec(): it is bound by latency/issue, not by the fp64 pipe or HBM (profiles/); both fractions are given so the distance to either roof is explicit",
    }

    line = {
        "metric": "expressions_evaluated_per_sec",
        "value": exprs_all / dev_all if dev_all > 0 else 0.0,
        "unit": "expr/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall_all / args.steps,
        "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload if world == 1 else (workload + ", ONE search over %d GPUs (K1 sharded by forms, NCCL all-gather)" % world) if sharded else "ries <t_rank> -l%s -F3, one independent target per GPU (BASELINE config-5 distribution, seed %d)" % (args.level, BATCH_SEED),
            "expressions_per_step": exprs_all / args.steps, "dead_ends_per_step": dead_all / args.steps,
            "l2": "every step rebuilds all device data from scratch; working set (table + survivors + nodes) exceeds the 126 MB L2 at this level",
            "value_is": "expressions / device-busy time (device clock stamps over all stages of every pass, cross-checked with CUDA events in roofline.cuda_event_check); e2e is the same search timed around the C-ABI call",
        },
        "time_to_full_report_s": wall_all / args.steps,
        "nodes_per_sec_e2e": (exprs_all + dead_all) / wall_all,
        "e2e": {"value": exprs_all / wall_all, "unit": "expr/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "ries_main(argc, argv, write_cb) via ctypes"},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roofline,
    }

    if sharded_extra is not None:
        line["sharded_search"] = sharded_extra
    if world == 1:
        # SURVEY 8(d) also defines time-to-full-report from PROCESS start: the `ries` binary in a fresh process
        # (CUDA context creation, cold allocations, search, teardown) -- what a command-line user sees; the
        # step time above is a warm call on a live engine, what the web worker's reused module sees
        exe = os.path.join(ROOT, "ries_b200", "bin", "ries")
        if os.path.exists(exe):
            cold = []
            for _ in range(2):
                tc0 = time.perf_counter()
                subprocess.run([exe, args.target, "-l" + args.level, "-F3", "--no-slow-messages"],
                               capture_output=True, env=dict(os.environ, RIES_DEVICE=str(local_rank)))
                cold.append(time.perf_counter() - tc0)
            line["time_to_full_report_cold_process_s"] = min(cold)
    if not args.no_cpu_baseline and world == 1:
        binary, kind = ref_binary()
        if binary:
            n, dt = run_cpu(binary, [args.target, "-l" + args.cpu_level, "-F3"])
            line["cpu_baseline"] = {"value": n / dt, "unit": "expr/s", "cores": 1, "kind": kind, "sample": sample_desc,
                                    "seconds": dt, "expressions": n}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
