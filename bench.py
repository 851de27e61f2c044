#!/usr/bin/env python3
"""bench.py -- expressions evaluated/sec and time-to-full-report of the RIES search on B200.

A "step" is one complete search (enumerate -> evaluate -> sort/dedupe/merge -> match/Newton -> report)
through the reference-facing C ABI (`ries_main(argc, argv, write_cb)`, include/ries_b200.h), i.e. exactly
the call the web worker makes (`callMain([...flags, target])`).  Nothing is reused between steps except
device allocations: every step rebuilds the whole value table from scratch.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
        N>1: ONE search over the N GPUs (strong scaling); every rank's report is compared with the reference's
    python bench.py --impl reference          (the reference's own CPU implementation, oracle/_ref/ries)

Workload (every N): `ries 2.5063141592653589 -l6 -F3`, the full-depth -l6 search BASELINE.json's metric is
quoted on (1.45e8 expressions, 2.04e9 dead-ends, 3.23e7 table rows; fits one GPU).  The N=1 line also carries
BASELINE configs 1, 2, 3, 4 (one timing each, report checked against the reference's committed output) and
config 5 (a batch of targets at -l4, with an 8-process CPU sample beside it).
"""
import argparse
import hashlib
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
# BASELINE.md config 5 distribution: f"{random.uniform(0,10):.6f}", seed 20261019
BATCH_SEED = 20261019

# ALGORITHMIC bytes.  K1 must write every complete expression once (32 B row + 2 B
# form index); the sort reads and writes each 32 B row once (64 B/key, whatever the number of digit passes);
# merge + neighbour search: read old + new, write merged (2 x 32 B) + one read of the merged table (32 B).
ROW_BYTES = 32
K1_BYTES_PER_EXPR = ROW_BYTES + 2
SORT_BYTES_PER_KEY = 2 * ROW_BYTES
MERGE_BYTES_PER_ROW = 3 * ROW_BYTES
# fp64 flops per exec() of each symbol: (value, extra for d/dx), SURVEY.md section 8(d)
FLOP_WEIGHTS = {"n": (1, 1), "s": (1, 2), "r": (1, 3), "q": (1, 2), "+": (7, 1), "-": (7, 1), "*": (1, 3), "/": (1, 5),
                "l": (25, 1), "E": (20, 1), "S": (50, 3), "C": (50, 3), "T": (45, 4), "^": (60, 30), "v": (62, 32),
                "L": (52, 8), "A": (45, 7)}

# the other BASELINE.json configs, timed once each on the N=1 line: (key, golden case name, argv)
SIDE_CONFIGS = [
    ("config1_default_level", "pi_l2_F3", ["3.14159265358979", "-F3"]),
    ("config2_l4", "t25063_l4_F3", ["2.5063", "-l4", "-F3"]),
    ("config3_phi_l6_early_exit", "phi_l6_early_exit", None),
    ("config4_i_l5", "l5i", ["143", "-i", "-l5", "-F3"]),
    ("config4_S_l5", "l5S", ["2.5063", "-l5", "-S1234+-*/x", "-F3"]),
]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def normalise(text, prog=None):
    text = re.sub(r"Time: [0-9.]+", "Time: X", text)
    return text.replace(prog, "ries") if prog else text


def golden_args(name):
    with open(os.path.join(ROOT, "tests", "golden", "cases.json")) as f:
        for c in json.load(f):
            if c["name"] == name:
                return c["args"]
    return None


def golden_text(name):
    p = os.path.join(ROOT, "tests", "golden", name + ".txt")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return f.read()


def shown_status(code):
    return code & 0xff if code != 0 else 0


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


# ------------------------------------------------------------------------------------- the CPU side
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ries")
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
    return parse_expressions(r.stdout), dt, r.stdout, r.returncode


def boot_id():
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            return f.read().strip()
    except OSError:
        return "unknown"


def cpu_cache_path(binary, argv):
    h = hashlib.sha1((binary + "\0" + "\0".join(argv)).encode()).hexdigest()[:16]
    return os.path.join("/tmp", "ries_b200_cpu_%s.json" % h)


def cpu_full_run(binary, kind, argv, allow_cache=True):
    """The reference's CPU implementation on the headline command line, run to completion ONCE on one host
    core (it is single-threaded).  The measurement is kept in /tmp keyed by (binary, argv) and this boot, so
    that the other bench invocations of the same session on the same box (N=2,4,8 of a scaling run; the B200
    arm's cpu_baseline) quote the same run instead of spending another 5-8 minutes each."""
    path = cpu_cache_path(binary, argv)
    if allow_cache and os.path.exists(path):
        try:
            with open(path) as f:
                c = json.load(f)
            if c.get("boot_id") == boot_id() and time.time() - c.get("when", 0) < 6 * 3600:
                c["cached"] = True
                return c
        except (OSError, ValueError):
            pass
    n, dt, text, rc = run_cpu(binary, argv)
    gname = "l6full" if argv[:3] == [DEFAULT_TARGET, "-l" + DEFAULT_LEVEL, "-F3"] else None
    gold = golden_text(gname) if gname else None
    c = {"expressions": n, "seconds": dt, "returncode": rc, "kind": kind, "argv": argv, "boot_id": boot_id(),
         "when": time.time(), "cached": False,
         "identical_to_golden": (normalise(text, binary) + "exit=%d\n" % rc == gold) if gold else None}
    try:
        with open(path, "w") as f:
            json.dump(c, f)
    except OSError:
        pass
    return c


def cpu_bounded_prefix(binary, argv, seconds):
    """A bounded sample of the SAME command: the reference runs with its own per-pass trace (-Dy) on a pty
    (line-buffered), is stopped after `seconds`, and the expressions it had generated at its last completed
    pass are divided by the time of that line.  (Expression rate falls with depth -- the tree grows -- so a
    prefix flatters the CPU; the full run of `--impl reference` is the same-config figure.)"""
    import pty
    env = dict(os.environ, HOME="/nonexistent")
    master, slave = pty.openpty()
    t0 = time.perf_counter()
    p = subprocess.Popen([binary] + argv + ["-Dy", "--no-slow-messages"], env=env, stdout=slave, stderr=subprocess.DEVNULL)
    os.close(slave)
    buf = b""
    best = (0.0, 0.0)
    done = False
    try:
        while True:
            now = time.perf_counter() - t0
            if now > seconds:
                break
            import select
            r, _, _ = select.select([master], [], [], max(0.05, min(1.0, seconds - now)))
            if not r:
                if p.poll() is not None:
                    done = True
                    break
                continue
            try:
                chunk = os.read(master, 65536)
            except OSError:
                done = True
                break
            if not chunk:
                done = True
                break
            buf += chunk
            stamp = time.perf_counter() - t0
            for m in re.finditer(rb"gen_total ([0-9.e+]+) \(", buf):
                best = (float(m.group(1)), stamp)
            buf = buf[-4096:]
    finally:
        if p.poll() is None:
            p.kill()
        p.wait()
        os.close(master)
    return {"expressions": best[0], "seconds": best[1], "ran_to_completion": done}


def targets_for(n):
    import random
    rng = random.Random(BATCH_SEED)
    out = []
    while len(out) < n:
        t = "%.6f" % rng.uniform(0, 10)
        if float(t) != 0.0:
            out.append(t)
    return out


def cpu_batch_sample(binary, level, n=8):
    """the reference has no batch mode -> n concurrent processes, one target each."""
    env = dict(os.environ, HOME="/nonexistent")
    ts = targets_for(n)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([binary, t, "-l" + level, "-F3", "--no-slow-messages"], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, text=True) for t in ts]
    exprs = sum(parse_expressions(p.communicate()[0]) for p in ps)
    dt = time.perf_counter() - t0
    return {"targets": n, "processes": n, "cores": min(n, os.cpu_count() or n), "seconds": dt,
            "targets_per_sec": n / dt, "expr_per_sec": exprs / dt,
            "sample": "the first %d targets of the batch, one reference process each, all at once" % n}


# ------------------------------------------------------------------------------------- B200 arm pieces
def batch_object(args, ries_b200, rank, world, dist, torch):
    """BASELINE config 5: targets ~ U(0,10) with 6 decimals (seed 20261019) at -l4 through ries_batch; the first
    64 reports are compared with the reference's (tests/golden/batch_l4.json, made by tools/make_batch_golden.py)."""
    from ries_b200.parallel import shard_targets
    level = args.batch_level
    alltargets = targets_for(args.batch_size * world)  # --batch-size is per GPU (weak scaling: independent units)
    mine = shard_targets(alltargets, rank, world)
    flags = ["-l" + level, "-F3", "--no-slow-messages"]
    ries_b200.batch(mine[:max(2 * args.workers, 8)], flags, workers=args.workers)  # warm-up: engines, graphs
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
    walls = []
    for _ in range(3):  # the median of three batches (run-to-run spread is large: 16 host threads share the driver)
        t0 = time.perf_counter()
        got = ries_b200.batch(mine, flags, workers=args.workers)
        walls.append(time.perf_counter() - t0)
        if dist is not None:
            dist.barrier()
    wall = sorted(walls)[1]
    exprs = sum(parse_expressions(text) for _, text in got)
    checked, same = 0, 0
    gpath = os.path.join(ROOT, "tests", "golden", "batch_l%s.json" % level)
    if os.path.exists(gpath):
        with open(gpath) as f:
            g = json.load(f)
        want = dict(zip(g["targets"], g["reports"]))
        for t, (code, text) in zip(mine, got):
            if t in want:
                checked += 1
                same += int(normalise(text) + "exit=%d\n" % shown_status(code) == want[t])
    nt = float(len(mine))
    if dist is not None:
        t = torch.tensor([exprs, nt, checked, same], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        tm = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        exprs, nt, checked, same = t.tolist()
        wall = tm.item()
    return {"workload": "BASELINE config 5: %d targets U(0,10) 6dp seed %d, `ries <t> -l%s -F3`, %d engines per GPU, "
                        "%d GPU(s), one RHS enumeration per GPU shared by its targets" % (args.batch_size * world, BATCH_SEED, level, args.workers, world),
            "targets": int(nt), "seconds": wall, "targets_per_sec": nt / wall, "targets_per_sec_per_gpu": nt / wall / world,
            "expr_per_sec": exprs / wall, "reports_checked_against_reference": int(checked),
            "reports_identical": int(same), "shared_rhs": ries_b200.batch_shared_stats(),
            "seconds_is": "median of three batches", "seconds_each_rank0": [round(w, 4) for w in walls],
            "api": "ries_batch(targets, flags, workers, done_cb) via ctypes"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--target", default=DEFAULT_TARGET)
    ap.add_argument("--level", default=DEFAULT_LEVEL)
    ap.add_argument("--cpu-sample-seconds", type=float, default=25.0,
                    help="length of the bounded CPU sample when no full reference run of this box is on record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the side configs and the batch object")
    ap.add_argument("--workload", default="search", choices=["search", "batch"],
                    help="search: one full-depth search per step (default); batch: BASELINE config 5 as the headline")
    ap.add_argument("--batch-size", type=int, default=256, help="targets per GPU (batch object / workload)")
    ap.add_argument("--batch-level", default="4")
    ap.add_argument("--workers", type=int, default=16, help="concurrent engines per GPU (batch)")
    ap.add_argument("--independent-targets", action="store_true",
                    help="N>1: one independent target per rank (weak scaling, no collective) instead of ONE search "
                         "over the N GPUs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    workload = "ries %s -l%s -F3" % (args.target, args.level)
    head_argv = [args.target, "-l" + args.level, "-F3"]
    is_headline = args.target == DEFAULT_TARGET and args.level == DEFAULT_LEVEL
    head_golden = golden_text("l6full") if is_headline else None

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        binary, kind = ref_binary()
        if binary is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ries and oracle/_build/ries_oracle are both missing"}))
            return 0
        c = cpu_full_run(binary, kind, head_argv)
        v = c["expressions"] / c["seconds"]
        sample = ("the whole search, run to completion once on one host core (the reference is single-threaded); "
                  "--steps/--warmup do not apply: steps_effective 1" + (", quoted from this box's earlier run of the same command" if c["cached"] else ""))
        line = {
            "impl": "reference", "metric": "expressions_evaluated_per_sec", "value": v, "unit": "expr/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "steps_effective": 1, "warmup_effective": 0,
            "ms_per_step": 1e3 * c["seconds"], "time_to_full_report_s": c["seconds"],
            "higher_is_better": True, "scaling": "strong" if n_gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "same_config_as_b200_arm": True, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "expr/s", "cores": 1, "kind": c["kind"], "sample": sample,
                             "seconds": c["seconds"], "expressions": c["expressions"], "cached": c["cached"],
                             "identical_to_golden": c.get("identical_to_golden")},
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

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "batch":
        b = batch_object(args, ries_b200, rank, world, dist, torch)
        if rank == 0:
            print(json.dumps({
                "metric": "expressions_evaluated_per_sec", "value": b["expr_per_sec"], "unit": "expr/s",
                "n_gpus": n_gpus, "steps": 1, "warmup": 1, "ms_per_step": 1e3 * b["seconds"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": b["workload"]}, "batch": b,
                "e2e": {"value": b["expr_per_sec"], "unit": "expr/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                        "api": b["api"]}}))
        if dist is not None:
            dist.destroy_process_group()
        return 0

    def join_search_communicator():
        """ries_dist_init on every rank: the id comes from rank 0 over torch.distributed (plumbing only)."""
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ries_b200.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ries_b200.dist_init(rank, world, uid.cpu().numpy().tobytes())

    one_search = world > 1 and not args.independent_targets
    if one_search:
        join_search_communicator()
        target = args.target
    elif world > 1:
        target = targets_for(world)[rank]  # independent targets: BASELINE config-5 style sharding
    else:
        target = args.target
    argv = [target, "-l" + args.level, "-F3", "--no-slow-messages"]

    def step(av=argv):
        t0 = time.perf_counter()
        code, text = ries_b200.call_main(av)
        dt = time.perf_counter() - t0
        st = ries_b200.last_stats()
        assert code in (0, -9998), (code, text[-400:])
        return dt, st, text, code

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

    # every timed step's report against the reference's (committed golden, made from the unmodified ries.c)
    identical = None
    if head_golden is not None and (world == 1 or one_search):
        ok = all(normalise(s[2]) + "exit=%d\n" % shown_status(s[3]) == head_golden for s in per_step)
        if dist is not None:
            f = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            ok = bool(f.item())
            h = int(hashlib.sha1(normalise(per_step[-1][2]).encode()).hexdigest()[:12], 16)
            hs = torch.tensor([h, -h], dtype=torch.int64, device="cuda")
            dist.all_reduce(hs, op=dist.ReduceOp.MAX)
            ok = ok and (hs[0].item() == -hs[1].item())  # max == min: every rank printed the same bytes
        identical = ok

    # Cross-check of the per-stage device times: the engine takes them from device clock stamps (%globaltimer)
    # that the first kernel of every stage leaves in the pass's control block; here a few extra steps, outside
    # the timed region, bracket every stage with CUDA events on the engine's stream instead (RIES_STAGE_EVENTS).
    event_check = None
    if world == 1:
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
    dev_total_s = sum(dev_ms) / 1e3
    last = per_step[-1][1]

    # whole-job numbers: sum of units over ranks, max of time over ranks
    if dist is not None:
        if one_search:  # every rank holds the totals of the one search: count them once
            exprs_c, dead_c = (exprs, dead) if rank == 0 else (0.0, 0.0)
        else:
            exprs_c, dead_c = exprs, dead
        t = torch.tensor([exprs_c, dead_c, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        tmax = torch.tensor([wall, dev_total_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        exprs_all, dead_all, launches_all = t.tolist()
        wall_all, dev_all = tmax.tolist()
    else:
        exprs_all, dead_all, launches_all, wall_all, dev_all = exprs, dead, launches, wall, dev_total_s

    # N>1: the same ranks also run independent targets (weak scaling, no collective) as a secondary figure
    independent = None
    if one_search:
        ries_b200.dist_finalize()
        if os.environ.get("RIES_BENCH_NO_INDEPENDENT") is None:
            iargv = [targets_for(world)[rank], "-l" + args.level, "-F3", "--no-slow-messages"]
            step(iargv)
            barrier()
            ti0 = time.perf_counter()
            isteps = [step(iargv) for _ in range(max(1, min(args.steps, 3)))]
            barrier()
            iwall = time.perf_counter() - ti0
            ie = sum(s[1]["lhs_expressions"] + s[1]["rhs_expressions"] for s in isteps)
            t = torch.tensor([ie], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            tm = torch.tensor([iwall], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            independent = {"workload": "ries <t_rank> -l%s -F3, one independent target per GPU (config-5 distribution), no collective" % args.level,
                           "scaling": "weak", "expr_per_s_e2e": t.item() / tm.item(), "steps": len(isteps),
                           "seconds_per_step": tm.item() / len(isteps)}
    if rank != 0:
        if dist is not None:
            if world > 1 and not args.no_extras and os.environ.get("RIES_BENCH_NO_BATCH") is None:
                batch_object(args, ries_b200, rank, world, dist, torch)
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_kind = peaks()
    # ---- roofline of the dominant kernel: K1 (k_expand), fp64-bound by SURVEY 8(d) ----
    # achieved = algorithmic flop (exec() calls per symbol, counted on the device, x SURVEY 8(d) weights) / K1's
    # device time; per launch = the same ratio (sum over launches / sum of launch durations)
    flops = 0.0
    execs = 0
    for st in (s[1] for s in per_step):
        for ch, c in st["exec_calls"].items():
            w, wdx = FLOP_WEIGHTS.get(ch, (0, 0))
            flops += c["rhs"] * w + c["lhs"] * (w + wdx)
            execs += c["rhs"] + c["lhs"]
    fp64_peak = ries_b200.fp64_peak_tflops()
    fp64_ach = flops / (dev_ms[0] / 1e3) / 1e12 if dev_ms[0] > 0 else 0.0
    traffic, traffic_src = None, None
    for tname in ("r2_l6_k_expand_traffic.json", "r1_l6_k_expand_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if is_headline and world == 1 and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"]
            traffic_src = ("profiles/%s: ncu dram__bytes_read.sum+dram__bytes_write.sum over the %d K1 launches of one "
                           "search (cold process), per launch" % (tname, tj["launches"]))
            break

    def hbm_stage(name, algo_bytes, ms, nlaunch):
        ach = algo_bytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"kernels": name, "algorithmic_bytes_per_step": algo_bytes / args.steps, "ms_per_step": ms / args.steps,
                "launches_per_step": nlaunch / args.steps, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak}
    k1_algo = exprs * K1_BYTES_PER_EXPR
    hbm = {
        "K1_output": hbm_stage("k_expand: 34 B per complete expression it must write", k1_algo, dev_ms[0], dev_launches[0]),
        "K2_sort": hbm_stage("k_keys + k_rs_onesweep: 64 B per key", exprs * SORT_BYTES_PER_KEY, dev_ms[1], dev_launches[1]),
        "K2K3_dedupe_match_merge": hbm_stage("k_classify_compact + k_match_merge_new + k_merge_old: 96 B per merged row",
                                             merge_rows * MERGE_BYTES_PER_ROW, dev_ms[2], dev_launches[2]),
        "peak_kind": peak_kind,
    }
    if traffic is not None and dev_launches[0]:
        hbm["K1_output"]["ncu_traffic_over_algorithmic"] = traffic * dev_launches[0] / max(k1_algo, 1.0)
    roofline = {
        "bound": "fp64", "kernel": "K1 enumerate+evaluate (k_expand)", "achieved": fp64_ach, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak if fp64_peak else None,
        "peak_kind": "measured here: DFMA microbenchmark on the engine's stream (MEASURED_PEAKS.json has no fp64 entry; SURVEY 8(d))",
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_flop_per_step": flops / args.steps, "exec_calls_per_step": execs / args.steps,
        "launches_per_step": dev_launches[0] / args.steps,
        "avg_launch_ms": dev_ms[0] / dev_launches[0] if dev_launches[0] else None,
        "stage_ms_per_step": {"K1_enum_eval": dev_ms[0] / args.steps, "K2_sort": dev_ms[1] / args.steps,
                              "K2K3_dedupe_match_merge": dev_ms[2] / args.steps, "K4_refine": dev_ms[3] / args.steps},
        "hbm": hbm,
        "sorted_keys_per_step": sort_keys / args.steps, "expressions_per_step": exprs / args.steps,
        "stage_timing": "device clock stamps (%globaltimer) written by the first kernel of each stage into the pass's control block, on the engine's stream, summed over the timed region",
        "cuda_event_check": event_check,
        "note": "K1 does ~14 algorithmic flop and ~10 integer tests per exec(): it is bound by instruction issue and latency, not by the fp64 pipe or HBM (profiles/); both fractions are given so the distance to either roof is explicit",
    }

    mode = ("ONE search over %d GPUs: (form, first symbol) pairs dealt round-robin for K1 + hash pre-dedupe, fixed-size "
            "NCCL all-gather of what survives, sort/dedupe/match split by value range, second small exchange, "
            "replicated merge; every rank prints the full report" % world)
    line = {
        "metric": "expressions_evaluated_per_sec",
        "value": exprs_all / wall_all if wall_all > 0 else 0.0,
        "unit": "expr/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall_all / args.steps,
        "higher_is_better": True, "scaling": "strong" if one_search else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload if world == 1 else (workload + ", " + mode) if one_search else "ries <t_rank> -l%s -F3, one independent target per GPU (BASELINE config-5 distribution, seed %d)" % (args.level, BATCH_SEED),
            "expressions_per_step": exprs_all / args.steps, "dead_ends_per_step": dead_all / args.steps,
            "l2": "every step rebuilds all device data from scratch; working set (table + survivors + nodes) exceeds the 126 MB L2 at this level",
            "value_is": "expressions / wall time of the timed region (barrier to barrier around the K calls of ries_main, max over ranks). The input of this path is a command line: there is nothing to pre-stage in HBM, so value and e2e are the same measurement; device_busy_expr_per_s is the kernel-only figure",
        },
        "time_to_full_report_s": wall_all / args.steps,
        "report_identical_to_reference": identical,
        "device_busy_expr_per_s": exprs_all / dev_all if dev_all > 0 else None,
        "nodes_per_sec_e2e": (exprs_all + dead_all) / wall_all,
        "e2e": {"value": exprs_all / wall_all, "unit": "expr/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "ries_main(argc, argv, write_cb) via ctypes"},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roofline,
    }
    if one_search:
        line["sharded_search"] = {
            "identical_on_every_rank": identical, "sharded_passes": last["sharded_passes"],
            "allgather_MB_per_search": last["gather_bytes"] / 1e6, "allgather_ms": last["gather_ms"],
            "stage_ms": last["stage_ms"], "how": mode}
    if independent is not None:
        line["independent_targets"] = independent

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
    if not args.no_extras:
        if world == 1:
            # the other BASELINE configs: warm call, best of 3, report compared with the reference's
            for key, gname, cargv in SIDE_CONFIGS:
                cargv = cargv or golden_args(gname)
                gold = golden_text(gname)
                if cargv is None or gold is None:
                    continue
                av = list(cargv) + ["--no-slow-messages"]
                best, same, st = None, True, None
                for _ in range(4):
                    t0 = time.perf_counter()
                    code, text = ries_b200.call_main(av)
                    dt = time.perf_counter() - t0
                    same = same and (normalise(text) + "exit=%d\n" % shown_status(code) == gold)
                    st = ries_b200.last_stats()
                    best = dt if best is None else min(best, dt)
                ne = st["lhs_expressions"] + st["rhs_expressions"]
                line[key] = {"argv": " ".join(cargv), "ms": 1e3 * best, "expressions": ne, "expr_per_s_e2e": ne / best,
                             "kernel_launches": st["kernel_launches"], "identical_to_reference": same}
        if os.environ.get("RIES_BENCH_NO_BATCH") is None:
            line["batch"] = batch_object(args, ries_b200, rank, world, dist, torch)
    if not args.no_cpu_baseline and world == 1:
        binary, kind = ref_binary()
        if binary:
            path = cpu_cache_path(binary, head_argv)
            if os.path.exists(path):
                c = cpu_full_run(binary, kind, head_argv)
                line["cpu_baseline"] = {"value": c["expressions"] / c["seconds"], "unit": "expr/s", "cores": 1, "kind": kind,
                                        "sample": "the whole search, run to completion on one host core of this box by `bench.py --impl reference` earlier in this session",
                                        "seconds": c["seconds"], "expressions": c["expressions"]}
            else:
                c = cpu_bounded_prefix(binary, head_argv, args.cpu_sample_seconds)
                if c["seconds"] > 0:
                    line["cpu_baseline"] = {"value": c["expressions"] / c["seconds"], "unit": "expr/s", "cores": 1, "kind": kind,
                                            "sample": "the first %.0f s of the same command (`%s -Dy`), expressions generated up to its last completed pass; "
                                                      "the rate falls with depth (the tree outgrows the caches), so the full run of `--impl reference` (8-9 min) is lower"
                                                      % (args.cpu_sample_seconds, workload),
                                            "seconds": c["seconds"], "expressions": c["expressions"]}
            if not args.no_extras and "batch" in line:
                line["batch"]["cpu_sample"] = cpu_batch_sample(binary, args.batch_level, 8)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if identical is False:
        print("bench.py: the search did not print the reference's report (on every rank)", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
